"""cuda_dom_BC_star (src/cuda_bluebottle.cu:2111-2311, BC_{u,v,w}_{W,E,S,N,B,T}_{D,N} src/bluebottle_kernel.cu:104-598),
SURVEY.md 8f rank 2.  CPU: the oracle restatement (bbo_dom_BC_star) against the golden outputs of the reference's own
kernels and against the properties the formulas encode.  GPU: the product (bbpcg_dom_BC_star / bbpcg_prologue, through the
C ABI) against the oracle, and live against the reference's kernels."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from cases import BC_STAR_TABLES, Case, face_exchange_inputs, load_ref, ref_dom_BC_star, velocity_bc
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST_PATH = os.path.join(GOLD, "BCSTAR_MANIFEST.json")
MANIFEST = json.load(open(MANIFEST_PATH)) if os.path.exists(MANIFEST_PATH) else None
STAR = (("u", ob.U_STAR, "u_star"), ("v", ob.V_STAR, "v_star"), ("w", ob.W_STAR, "w_star"))
TOL = 4e-15          # the reference kernels and the product are FMA-contracted by nvcc, the oracle is not


def _seed(case, seed):
    for r in range(case.o.nblocks):
        fx = face_exchange_inputs(case, r, seed)
        for key, aid, _ in STAR:
            case.o.array(r, aid)[...] = fx[key][0]


def _axis_view(a, comp, axis):
    """view of component comp's face-grid array (Gfx a[i,k,j], Gfy a[j,i,k], Gfz a[k,j,i]) with the index of `axis`
    (0 x, 1 y, 2 z) first"""
    from bbpcg.grid import as_ijk
    return np.moveaxis(as_ijk(a, ("Gfx", "Gfy", "Gfz")[comp]), axis, 0)


# ---- CPU: the oracle ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("table", ["dirichlet", "neumann", "mixed"])
def test_oracle_bc_star_encodes_the_wall_values(table):
    """DIRICHLET: the wall value is bc -- exactly on the wall face for the normal component (and the ghost mirrors the
    first inner face about it), as the quadratic through ghost / first / second evaluated at the wall for a tangential one:
    3/8 ghost + 3/4 first - 1/8 second = bc.  NEUMANN: ghost = first.  PERIODIC / PRECURSOR: untouched."""
    case = Case((7, 6, 5), bc="box")
    _seed(case, 5)
    before = {k: case.o.array(0, aid).copy() for k, aid, _ in STAR}
    types, vals = BC_STAR_TABLES[table]
    case.o.dom_BC_star(types, vals)
    for c, (k, aid, _) in enumerate(STAR):
        a = case.o.array(0, aid)
        changed = a != before[k]
        allowed = np.zeros_like(changed)
        for f in range(6):
            axis, high = f // 2, f & 1
            ty, bc = types[c * 6 + f], vals[c * 6 + f]
            v, v0 = _axis_view(a, c, axis), _axis_view(before[k], c, axis)
            al = _axis_view(allowed, c, axis)
            g, first, second = (-1, -2, -3) if high else (0, 1, 2)
            core = (slice(1, -1), slice(1, -1))
            if ty in (1, 2):
                al[g][core] = True
            if ty == 1 and c == axis:
                al[first][core] = True
                assert np.all(v[first][core] == bc)
                inner = (slice(2, -2), slice(2, -2))
                assert np.abs(0.5 * (v[g][inner] + v0[second][inner]) - bc).max() <= 1e-14 * max(1., abs(bc))
            elif ty == 1:
                inner = (slice(2, -2), slice(2, -2))            # away from the wall faces that LATER faces overwrite
                wall = 0.375 * v[g][inner] + 0.75 * v[first][inner] - 0.125 * v[second][inner]
                assert np.abs(wall - bc).max() <= 1e-13 * (abs(bc) + np.abs(v[first][inner]).max())
            elif ty == 2:
                inner = (slice(2, -2), slice(2, -2))
                assert np.array_equal(v[g][inner], v[first][inner])
        assert not (changed & ~allowed).any(), k             # faces only: edges, corners and untouched entries keep their bits


def test_oracle_bc_star_skips_sides_with_a_neighbour():
    """`dom[rank].w == MPI_PROC_NULL` (cuda_bluebottle.cu:2114): internal block boundaries and periodic wraps get no BC"""
    case = Case((8, 6, 4), blocks=(2, 1, 2), bc="duct")      # x periodic (neighbour = the other block), y / z walls
    _seed(case, 9)
    before = [[case.o.array(r, aid).copy() for _, aid, _ in STAR] for r in range(case.o.nblocks)]
    case.o.dom_BC_star(*BC_STAR_TABLES["dirichlet"])
    for r in range(case.o.nblocks):
        d = case.o.dom(r)
        for c, (_, aid, _) in enumerate(STAR):
            diff = case.o.array(r, aid) != before[r][c]
            assert not _axis_view(diff, c, 0)[0].any() and not _axis_view(diff, c, 0)[-1].any()       # W / E: neighbours everywhere
            zlo, zhi = _axis_view(diff, c, 2)[0], _axis_view(diff, c, 2)[-1]
            assert zlo.any() == (d.b < 0) and zhi.any() == (d.t < 0)


@pytest.mark.skipif(MANIFEST is None, reason="tests/golden/bcs_*.npz not generated yet (oracle/make_golden_bcstar.py, needs a GPU)")
@pytest.mark.parametrize("name", sorted(MANIFEST["cases"]) if MANIFEST else [])
def test_oracle_bc_star_matches_reference_kernels(name):
    spec = MANIFEST["cases"][name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    for t, (types, vals) in MANIFEST["tables"].items():
        case = Case(tuple(spec["cells"]), bc=spec["bc"])
        _seed(case, MANIFEST["seed"])
        fresh = {k: v[0] for k, v in face_exchange_inputs(case, 0, MANIFEST["seed"]).items()}
        case.o.dom_BC_star(types, vals)
        for k, aid, _ in STAR:
            mine, ref = case.o.array(0, aid), gold["%s_%s" % (t, k)]
            assert np.array_equal(mine != fresh[k], ref != fresh[k]) or np.abs(mine - ref).max() <= TOL * np.abs(ref).max(), (t, k)
            assert np.abs(mine - ref).max() <= TOL * np.abs(ref).max(), (t, k)


# ---- GPU: the product ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cells,blocks,bc", [((24, 18, 16), (1, 1, 1), "box"), ((2, 1, 3), (1, 1, 1), "cavity"), ((33, 20, 17), (1, 1, 1), "duct"),
                                              ((24, 18, 16), (2, 1, 2), "box"), ((12, 18, 8), (1, 3, 1), "cavity"), ((16, 16, 16), (2, 2, 2), "channel")])
@pytest.mark.parametrize("table", ["dirichlet", "neumann", "cavity_lid", "mixed"])
def test_bc_star_matches_oracle_and_reference(cells, blocks, bc, table):
    from gpu_util import Product
    case = Case(cells, blocks=blocks, bc=bc)
    p = Product(case)
    _seed(case, 13)
    for r in range(p.n):
        for _, aid, name in STAR:
            p.dev[r][name] = p.solvers[r].to_device(case.o.array(r, aid))
    vbc = velocity_bc(table)
    case.o.dom_BC_star(*BC_STAR_TABLES[table])
    p.each(lambda r, s, d: s.dom_BC_star(d["u_star"], d["v_star"], d["w_star"], vbc))
    for r in range(p.n):
        fresh = face_exchange_inputs(case, r, 13)
        for k, aid, name in STAR:
            mine, ref = p.dev[r][name].cpu().numpy(), case.o.array(r, aid)
            same = ref == fresh[k][0]
            assert np.array_equal(mine[same], ref[same]), (r, name)                 # what the reference leaves alone keeps its bits
            assert np.abs(mine - ref).max() <= TOL * max(1., np.abs(ref).max()), (r, name)
    if blocks == (1, 1, 1):
        lib = load_ref()
        if lib is not None:
            dom, DOM = case.o.dom(0), case.o.DOM
            assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
            arrs = {k: np.ascontiguousarray(v[0]).copy() for k, v in face_exchange_inputs(case, 0, 13).items()}
            ref_dom_BC_star(lib, case, arrs, table)
            for k, _, name in STAR:
                assert np.abs(p.dev[0][name].cpu().numpy() - arrs[k]).max() <= TOL * max(1., np.abs(arrs[k]).max()), k
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("blocks,out_plane", [((1, 1, 1), "HOMOGENEOUS"), ((2, 1, 2), "EAST"), ((1, 2, 2), "TOP")])
def test_prologue_equals_the_reference_sequence(blocks, out_plane):
    """bbpcg_prologue = src/bluebottle.c:213-225 without particles: BC_star, three exchanges, solvability, BC_star, three
    exchanges -- against the same sequence on the oracle"""
    from bbpcg.lib import OUT_PLANE
    from gpu_util import Product
    case = Case((24, 18, 16), blocks=blocks, bc="cavity")
    p = Product(case)
    _seed(case, 29)
    scale = 0.0
    for r in range(p.n):
        for _, aid, name in STAR:
            p.dev[r][name] = p.solvers[r].to_device(case.o.array(r, aid))
            scale += np.abs(case.o.array(r, aid)).sum() * 0.01
    types, vals = BC_STAR_TABLES["mixed"]
    for step in range(2):
        if step == 1:
            case.o.solvability(OUT_PLANE[out_plane])
        case.o.dom_BC_star(types, vals)
        for _, aid, _ in STAR:
            case.o.exchange(aid)
    vbc = velocity_bc("mixed")
    ms = p.each(lambda r, s, d: s.prologue(d["u_star"], d["v_star"], d["w_star"], vbc, out_plane))
    assert all(m > 0. for m in ms)
    for r in range(p.n):
        for _, aid, name in STAR:
            mine, ref = p.dev[r][name].cpu().numpy(), case.o.array(r, aid)
            assert np.abs(mine - ref).max() <= 1e-12 * max(1., np.abs(ref).max()) + 1e-12 * scale / max(1, ref.size) , (r, name)
    p.close()
