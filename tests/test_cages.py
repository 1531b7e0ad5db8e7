"""Coefficient producers: cuda_build_cages (src/cuda_particle.cu:1516-1646, kernels
src/particle_kernel.cu:79-576).  CPU: the oracle restatement (bbo_build_cages) against the golden outputs of the reference's
own kernels.  GPU: bbpcg_build_cages (one CTA per particle + one fused flag / mask pass, through the C ABI) against the
oracle and, live, against the reference's kernels -- and the solve that follows it WITHOUT bbpcg_set_coefficients."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from cases import Case, cage_particles, load_ref, ref_build_cages, rel_l2
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST_PATH = os.path.join(GOLD, "CAGES_MANIFEST.json")
MANIFEST = json.load(open(MANIFEST_PATH)) if os.path.exists(MANIFEST_PATH) else None
INT_ARRAYS = (("flag_u", ob.FLAG_U), ("flag_v", ob.FLAG_V), ("flag_w", ob.FLAG_W), ("phase", ob.PHASE), ("phase_shell", ob.PHASE_SHELL))


def _oracle_case(cells, bc, parts_name, blocks=(1, 1, 1)):
    case = Case(cells, blocks=blocks, bc=bc)
    parts = cage_particles(parts_name, case.extent, case.cells) if parts_name else None
    if parts:
        case.o.build_cages(*parts)
    else:
        case.o.build_flags_noparts()
    case.o.jacobi_init()                                  # invM follows the flags just built
    return case, parts


# ---- CPU ----------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(MANIFEST is None, reason="tests/golden/cage_*.npz not generated yet (oracle/make_golden_cages.py, needs a GPU)")
@pytest.mark.parametrize("name", sorted(MANIFEST["cases"]) if MANIFEST else [])
def test_oracle_cages_match_reference_kernels(name):
    spec = MANIFEST["cases"][name]
    case, _ = _oracle_case(tuple(spec["cells"]), spec["bc"], spec["parts"])
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    for k, aid in INT_ARRAYS:
        if k in gold.files:
            assert np.array_equal(case.o.array(0, aid), gold[k]), k


def test_oracle_cage_semantics():
    """phase = index of the LAST particle whose sphere holds the cell centre; the shell is the solid layer with a fluid
    stencil neighbour; a face flag is -1 exactly between solid and fluid or inside the shell"""
    case, parts = _oracle_case((24, 20, 28), "sedimentation", "inside")
    px, py, pz, pr = parts
    d = case.o.dom(0)
    ph, sh = case.o.array(0, ob.PHASE), case.o.array(0, ob.PHASE_SHELL)
    k, j, i = np.meshgrid(np.arange(ph.shape[0]), np.arange(ph.shape[1]), np.arange(ph.shape[2]), indexing="ij")
    x, y, z = (i - 0.5) * d.dx + d.xs, (j - 0.5) * d.dy + d.ys, (k - 0.5) * d.dz + d.zs
    expect = np.full(ph.shape, -1)
    for n in range(len(px)):
        dist = np.sqrt((x - px[n]) ** 2 + (y - py[n]) ** 2 + (z - pz[n]) ** 2)
        expect[dist / pr[n] < 1 - 1e-9] = n
    inner = (slice(2, -2),) * 3                           # away from the z walls, where the cage is clipped
    far = np.ones(ph.shape, bool)
    for n in range(len(px)):                              # skip cells within round-off of a sphere surface
        dist = np.sqrt((x - px[n]) ** 2 + (y - py[n]) ** 2 + (z - pz[n]) ** 2)
        far &= np.abs(dist / pr[n] - 1) > 1e-9
    assert np.array_equal(ph[inner][far[inner]], expect[inner][far[inner]])
    assert (ph == 2).any() and (ph == 1).any()            # the overlapping pair: both own cells, the later one the shared ones
    assert set(np.unique(sh)) == {0, 1} and not (sh[ph < 0] == 0).any()
    fu = case.o.array(0, ob.FLAG_U)                       # a[i, k, j]
    assert set(np.unique(fu)) <= {-1, 0, 1}


# ---- GPU ----------------------------------------------------------------------------------------------------------------
def _run_product(case, parts, NPARTS=None):
    import torch
    from gpu_util import Product
    p = Product(case)
    xyzr = None
    if parts is not None and len(parts[0]):
        xyzr = np.stack(parts, axis=1)
    outs = []
    for r in range(p.n):
        s, d = p.solvers[r], p.dev[r]
        for k in ("flag_u", "flag_v", "flag_w", "phase", "phase_shell"):
            d[k] = torch.full_like(d[k], 7)               # garbage: every entry must be written
        d["parts"] = None if xyzr is None else s.to_device(xyzr)
    p.each(lambda r, s, d: s.build_cages(d["parts"], d["flag_u"], d["flag_v"], d["flag_w"], d["phase"], d["phase_shell"], NPARTS=NPARTS))
    for r in range(p.n):
        outs.append({k: p.dev[r][k].cpu().numpy() for k in ("flag_u", "flag_v", "flag_w", "phase", "phase_shell")})
    return p, outs


@pytest.mark.gpu
@pytest.mark.parametrize("cells,bc,parts_name,blocks", [
    ((24, 20, 28), "sedimentation", "inside", (1, 1, 1)), ((24, 20, 28), "sedimentation", "faces", (1, 1, 1)),
    ((21, 17, 19), "box", "faces", (1, 1, 1)), ((16, 16, 16), "periodic", "inside", (1, 1, 1)), ((40, 24, 36), "duct", "single", (1, 1, 1)),
    ((24, 20, 28), "sedimentation", "faces", (2, 1, 2)), ((24, 24, 24), "cavity", "inside", (2, 2, 2)), ((24, 20, 28), "channel", "faces", (1, 2, 1))])
def test_build_cages_matches_oracle_and_reference(cells, bc, parts_name, blocks):
    case, parts = _oracle_case(cells, bc, parts_name, blocks)
    p, outs = _run_product(case, parts)
    for r in range(p.n):
        for k, aid in INT_ARRAYS:
            assert np.array_equal(outs[r][k], case.o.array(r, aid)), (r, k)
    if blocks == (1, 1, 1):
        lib = load_ref()
        if lib is not None:
            dom, DOM = case.o.dom(0), case.o.DOM
            assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
            ref = ref_build_cages(lib, case, parts)
            for k, _ in INT_ARRAYS:
                assert np.array_equal(outs[0][k], ref[k]), k
    # the solve right after it, with NO bbpcg_set_coefficients: the masks were written by the same pass
    res = p.each(lambda r, s, d: s.PP_cg(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], d["phase"], d["phase_shell"]))[0]
    ores, _ = case.o.solve(parts=True)
    assert res.status == "converged" and abs(res.niter - ores.niter) <= 1
    assert rel_l2(p.gather("phi"), case.o.gather_interior(ob.PHI)) < 1e-10
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("bc,blocks", [("duct", (1, 1, 1)), ("cavity", (2, 1, 2)), ("periodic", (1, 1, 1))])
def test_build_cages_without_particles_writes_wall_flags_only(bc, blocks):
    """NPARTS == 0 (src/cuda_particle.cu:1524): flags = 1, 0 on external walls; phase arrays are not touched"""
    case, _ = _oracle_case((20, 12, 16), bc, None, blocks)
    p, outs = _run_product(case, None, NPARTS=0)
    for r in range(p.n):
        for k, aid in INT_ARRAYS[:3]:
            assert np.array_equal(outs[r][k], case.o.array(r, aid)), (r, k)
        assert (outs[r]["phase"] == 7).all() and (outs[r]["phase_shell"] == 7).all()
    res = p.each(lambda r, s, d: s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"]))[0]
    ores, _ = case.o.solve()
    assert res.status == "converged" and res.niter == ores.niter
    assert rel_l2(p.gather("phi"), case.o.gather_interior(ob.PHI)) < 1e-10
    p.close()
