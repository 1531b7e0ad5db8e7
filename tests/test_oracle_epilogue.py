"""CPU tests (no GPU) of the solve-epilogue restatement in the oracle (oracle/pcg_ref.c: bbo_dom_BC_p, bbo_project,
bbo_update_p = cuda_dom_BC_p / cuda_project / cuda_update_p, src/cuda_bluebottle.cu:2495-2589).

tests/golden/epi_*.npz are OUTPUTS OF THE REFERENCE'S OWN KERNELS (oracle/make_golden_epilogue.py, run on a B200).
Tolerances: ghost fills are copies -> bit-exact; u, v, w agree to the FMA contraction nvcc applies to
`u_star - dt/rho_f*gradPhi` (1e-14 of the field's max); p agrees to the summation order of the mean (1e-12).
Also: properties that need no golden vector (decomposition independence, zero mean, untouched edges)."""
import json
import os

import numpy as np
import pytest

from cases import Case
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = os.path.join(GOLD, "EPILOGUE_MANIFEST.json")
MANIFEST = json.load(open(MAN)) if os.path.exists(MAN) else {"cases": {}, "seed": 23}
CASES = MANIFEST["cases"]

VEL_TOL = 1e-14
P_TOL = 1e-12


def _gather_faces(case, aid, axis):
    """global face field from the blocks' Gf? arrays; the face shared by two blocks must agree"""
    o = case.o
    D = o.DOM
    shape = {0: (D.xn + 1, D.zn, D.yn), 1: (D.yn + 1, D.xn, D.zn), 2: (D.zn + 1, D.yn, D.xn)}[axis]
    out = np.full(shape, np.nan)
    for r in range(o.nblocks):
        d = o.dom(r)
        a = o.array(r, aid)[1:-1, 1:-1, 1:-1]
        i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
        if axis == 0:
            sl = (slice(i0, i0 + d.xn + 1), slice(k0, k0 + d.zn), slice(j0, j0 + d.yn))
        elif axis == 1:
            sl = (slice(j0, j0 + d.yn + 1), slice(i0, i0 + d.xn), slice(k0, k0 + d.zn))
        else:
            sl = (slice(k0, k0 + d.zn + 1), slice(j0, j0 + d.yn), slice(i0, i0 + d.xn))
        old = out[sl]
        both = ~np.isnan(old)
        assert np.array_equal(old[both], a[both])       # duplicated block-boundary faces carry the same value
        out[sl] = a
    assert not np.isnan(out).any()
    return out


def test_manifest_lists_every_epilogue_fixture():
    assert CASES and sorted(CASES) == sorted(f[:-4] for f in os.listdir(GOLD) if f.startswith("epi_") and f.endswith(".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_epilogue_matches_reference_kernels(name):
    spec = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    case = Case(tuple(spec["cells"]), bc=spec["bc"], nparts=spec.get("nparts", 0), radius=spec.get("radius", 1.0))
    case.seed_epilogue(MANIFEST["seed"])
    ein = case.epilogue_inputs(0)
    assert float(np.abs(ein["phi"]).sum() + np.abs(ein["p0"]).sum()) == float(gold["input_checksum"])
    case.o.epilogue(MANIFEST["rho_f"], MANIFEST["dt"])
    assert np.array_equal(case.o.array(0, ob.PHI), gold["phi"])               # exchange + Neumann copy: pure copies
    for key, aid in (("u", ob.U), ("v", ob.V), ("w", ob.W)):
        mine, ref = case.o.array(0, aid)[1:-1, 1:-1, 1:-1], gold[key][1:-1, 1:-1, 1:-1]
        assert np.abs(mine - ref).max() <= VEL_TOL * np.abs(ref).max(), key
    mine, ref = case.o.array(0, ob.P)[1:-1, 1:-1, 1:-1], gold["p"][1:-1, 1:-1, 1:-1]
    assert np.abs(mine - ref).max() <= P_TOL * np.abs(ref).max()


@pytest.mark.parametrize("bc", ["cavity", "duct", "box", "periodic"])
def test_dom_BC_p_touches_only_wall_faces(bc):
    from bbpcg.grid import BC_SETS, NEUMANN
    case = Case((9, 7, 8), bc=bc)
    case.seed_epilogue(5)
    before = case.o.array(0, ob.PHI).copy()
    case.o.dom_BC_p(ob.PHI)
    after = case.o.array(0, ob.PHI)
    pW, pE, pS, pN, pB, pT = BC_SETS[bc]
    I = slice(1, -1)
    faces = {"W": ((I, I, 0), (I, I, 1), pW), "E": ((I, I, -1), (I, I, -2), pE), "S": ((I, 0, I), (I, 1, I), pS),
             "N": ((I, -1, I), (I, -2, I), pN), "B": ((0, I, I), (1, I, I), pB), "T": ((-1, I, I), (-2, I, I), pT)}
    expect = before.copy()
    for ghost, inner, t in faces.values():
        if t == NEUMANN:                       # single block: a NEUMANN side has no neighbour
            expect[ghost] = before[inner]
    assert np.array_equal(after, expect)       # edges, corners and periodic sides untouched


@pytest.mark.parametrize("blocks,bc", [((2, 1, 1), "duct"), ((1, 2, 2), "cavity"), ((2, 2, 2), "periodic"), ((1, 1, 3), "sedimentation")])
def test_epilogue_is_decomposition_independent(blocks, bc):
    """same global phi/p0 on 1 block and on a decomposition: identical u, v, w (bit for bit) and p (to the mean's summation order)"""
    cells = (12, 10, 12)
    one, many = Case(cells, bc=bc), Case(cells, blocks=blocks, bc=bc)
    rng = np.random.default_rng(3)
    gphi, gp0 = rng.standard_normal(cells[::-1]), rng.standard_normal(cells[::-1])
    for case in (one, many):
        for r in range(case.o.nblocks):
            d = case.o.dom(r)
            i0, j0, k0 = d.Gcc.get("is") - 1, d.Gcc.get("js") - 1, d.Gcc.get("ks") - 1
            sl = (slice(k0, k0 + d.zn), slice(j0, j0 + d.yn), slice(i0, i0 + d.xn))
            case.o.array(r, ob.PHI)[...] = 7.0                              # ghosts start wrong everywhere
            case.o.array(r, ob.PHI)[1:-1, 1:-1, 1:-1] = gphi[sl]
            case.o.array(r, ob.P0)[1:-1, 1:-1, 1:-1] = gp0[sl]
        case.o.epilogue(1.0, 1e-3)
    for aid, axis in ((ob.U, 0), (ob.V, 1), (ob.W, 2)):
        assert np.array_equal(_gather_faces(one, aid, axis), _gather_faces(many, aid, axis))
    p1, pm = one.o.gather_interior(ob.P), many.o.gather_interior(ob.P)
    assert np.abs(p1 - pm).max() <= 1e-13 * np.abs(p1).max()
    assert abs(pm.mean()) <= 1e-13 * np.abs(pm).max()                          # mean pressure removed


def test_update_p_zeroes_solid_cells_before_the_mean():
    """p = (phase < 0)(p0 + phi) - mean: solid cells end at exactly -mean (src/bluebottle_kernel.cu:2396, :1507-1518)"""
    case = Case((16, 16, 16), bc="sedimentation", nparts=1, radius=3.0)
    case.seed_epilogue(9)
    mean = case.o.update_p()
    phase = case.o.array(0, ob.PHASE)[1:-1, 1:-1, 1:-1]
    p = case.o.array(0, ob.P)[1:-1, 1:-1, 1:-1]
    assert (phase > -1).sum() > 0
    assert np.array_equal(p[phase > -1], np.full((phase > -1).sum(), -mean))
    fluid = phase < 0
    expect = (case.o.array(0, ob.P0) + case.o.array(0, ob.PHI))[1:-1, 1:-1, 1:-1]
    assert np.array_equal(p[fluid], expect[fluid] + (-mean))


def test_projection_removes_the_divergence():
    """after a converged solve, div(u) = div(u*) - dt/rho L(phi) is the solver's residual: the projected field is
    discretely divergence free to the solve tolerance (what the pressure-Poisson step is for)"""
    case = Case((24, 20, 16), bc="duct")
    res, _ = case.solve_oracle(pp_residual=1e-10)
    assert res.status == 0
    case.seed_epilogue(1, phi=False)
    case.o.epilogue(1.0, 1e-3)
    d = case.o.dom(0)

    def div(u, v, w):
        u, v, w = u[1:-1, 1:-1, 1:-1], v[1:-1, 1:-1, 1:-1], w[1:-1, 1:-1, 1:-1]
        du = (u[1:] - u[:-1]).transpose(1, 2, 0)          # Gfx a[i,k,j] -> [k,j,i]
        dv = (v[1:] - v[:-1]).transpose(2, 0, 1)          # Gfy a[j,i,k] -> [k,j,i]
        dw = w[1:] - w[:-1]
        return du / d.dx + dv / d.dy + dw / d.dz
    a = case.o.array
    before = div(a(0, ob.U_STAR), a(0, ob.V_STAR), a(0, ob.W_STAR))
    after = div(a(0, ob.U), a(0, ob.V), a(0, ob.W))
    assert np.linalg.norm(after) < 1e-6 * np.linalg.norm(before)


# ---- face-grid halo exchanges: mpi_cuda_exchange_Gfx / _Gfy / _Gfz (src/mpi_comm.c:317-405) ---------------------
FACE_IDS = {"u": ob.U, "v": ob.V, "w": ob.W}


@pytest.mark.parametrize("name", sorted(CASES))
def test_face_exchange_matches_reference_kernels(name):
    """bit-exact against the reference's pack / unpack kernels (golden vectors, single block: periodic self-wrap)"""
    from cases import face_exchange_inputs
    spec = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    if "ex_u" not in gold.files:
        pytest.skip("fixture predates the face-exchange vectors")
    case = Case(tuple(spec["cells"]), bc=spec["bc"], nparts=spec.get("nparts", 0), radius=spec.get("radius", 1.0))
    for key, (arr, _) in face_exchange_inputs(case, 0, MANIFEST["seed"] + 6).items():
        case.o.array(0, FACE_IDS[key])[...] = arr
        case.o.exchange(FACE_IDS[key])
        assert np.array_equal(case.o.array(0, FACE_IDS[key]), gold["ex_" + key]), key


@pytest.mark.parametrize("blocks", [(1, 1, 1), (2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 2), (3, 1, 2)])
def test_face_exchange_reproduces_a_periodic_field_in_the_ghosts(blocks):
    """the reference's own halo test idea (cuda_BC_test_periodic, src/cuda_testing.cu:749-933) on all four grids: a smooth
    periodic function of the GLOBAL position, written to the interior only, must appear in every ghost FACE after the
    exchange -- across block boundaries and across the periodic wrap, with the shared face of a face grid skipped"""
    cells = (12, 8, 12)
    case = Case(cells, blocks=blocks, bc="periodic")
    o = case.o
    D = o.DOM
    L = (D.xe - D.xs, D.ye - D.ys, D.ze - D.zs)

    def field(x, y, z):
        return np.cos(2 * np.pi * x / L[0]) + 2 * np.cos(2 * np.pi * y / L[1]) + 3 * np.cos(2 * np.pi * z / L[2])
    for aid, grid in ((ob.PB_Q, "Gcc"), (ob.U, "Gfx"), (ob.V, "Gfy"), (ob.W, "Gfz")):
        exact = []
        for r in range(o.nblocks):
            d = o.dom(r)
            n = [d.xn + (grid == "Gfx"), d.yn + (grid == "Gfy"), d.zn + (grid == "Gfz")]
            # position of local index l (ghosts 0 and n+1): cell centres (l - 0.5) h, faces (l - 1) h, from the block's start
            ax = []
            for a_, (s0, h, nn, face) in enumerate(((d.xs, d.dx, n[0], grid == "Gfx"), (d.ys, d.dy, n[1], grid == "Gfy"), (d.zs, d.dz, n[2], grid == "Gfz"))):
                l = np.arange(nn + 2)
                ax.append(s0 + ((l - 1.0) if face else (l - 0.5)) * h)
            X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")            # [i, j, k]
            F = field(X, Y, Z)
            full = {"Gcc": F.transpose(2, 1, 0), "Gfz": F.transpose(2, 1, 0), "Gfx": F.transpose(0, 2, 1), "Gfy": F.transpose(1, 0, 2)}[grid]
            exact.append(np.ascontiguousarray(full))
            a = o.array(r, aid)
            a[...] = -99.0
            a[1:-1, 1:-1, 1:-1] = full[1:-1, 1:-1, 1:-1]
        o.exchange(aid)
        for r in range(o.nblocks):
            a, ex = o.array(r, aid), exact[r]
            I = slice(1, -1)
            for sl in ((0, I, I), (-1, I, I), (I, 0, I), (I, -1, I), (I, I, 0), (I, I, -1)):
                assert np.abs(a[sl] - ex[sl]).max() < 1e-12, (grid, r, sl)
            # edges and corners are NOT exchanged (faces only)
            assert a[0, 0, 0] == -99.0 and a[-1, -1, 0] == -99.0 and a[0, -1, 5] == -99.0


# ---- cuda_solvability (src/cuda_bluebottle.cu:2313-2492) ----------------------------------------------------------
def _seed_star(case, seed):
    from cases import face_exchange_inputs
    for r in range(case.o.nblocks):
        fx = face_exchange_inputs(case, r, seed)
        for key, aid in (("u", ob.U_STAR), ("v", ob.V_STAR), ("w", ob.W_STAR)):
            case.o.array(r, aid)[...] = fx[key][0]


def _net_flux(case):
    """sum over the global boundary faces of u* . n dA, from the blocks that touch them"""
    o, D = case.o, case.o.DOM
    tot = 0.0
    for r in range(o.nblocks):
        d = o.dom(r)
        u, v, w = (o.array(r, a)[1:-1, 1:-1, 1:-1] for a in (ob.U_STAR, ob.V_STAR, ob.W_STAR))
        if d.I == D.In - 1:
            tot += u[-1].sum() * d.dy * d.dz
        if d.I == 0:
            tot -= u[0].sum() * d.dy * d.dz
        if d.J == D.Jn - 1:
            tot += v[-1].sum() * d.dz * d.dx
        if d.J == 0:
            tot -= v[0].sum() * d.dz * d.dx
        if d.K == D.Kn - 1:
            tot += w[-1].sum() * d.dx * d.dy
        if d.K == 0:
            tot -= w[0].sum() * d.dx * d.dy
    return tot


@pytest.mark.parametrize("blocks", [(1, 1, 1), (2, 1, 2), (1, 3, 1)])
@pytest.mark.parametrize("out_plane", [10, 0, 1, 2, 3, 4, 5])
def test_solvability_removes_the_net_boundary_flux(blocks, out_plane):
    """the property the function exists for: afterwards the net outflow through the global boundary is zero (to round-off),
    only the designated plane(s) changed, and by a constant"""
    case = Case((12, 9, 8), blocks=blocks, bc="box")
    _seed_star(case, 17)
    before = [[case.o.array(r, a).copy() for a in (ob.U_STAR, ob.V_STAR, ob.W_STAR)] for r in range(case.o.nblocks)]
    flux0 = _net_flux(case)
    eps = case.o.solvability(out_plane)
    assert abs(sum(eps) - flux0) <= 1e-12 * abs(flux0)
    scale = sum(np.abs(b).sum() for blk in before for b in blk) * 0.01
    assert abs(_net_flux(case)) <= 1e-13 * scale
    D = case.o.DOM
    for r in range(case.o.nblocks):
        d = case.o.dom(r)
        for g, aid in enumerate((ob.U_STAR, ob.V_STAR, ob.W_STAR)):
            diff = case.o.array(r, aid) - before[r][g]
            lo_rank = (d.I, d.J, d.K)[g] == 0
            hi_rank = (d.I, d.J, d.K)[g] == (D.In, D.Jn, D.Kn)[g] - 1
            touched_lo = (out_plane == 10 or out_plane == 2 * g) and lo_rank
            touched_hi = (out_plane == 10 or out_plane == 2 * g + 1) and hi_rank
            inner = diff[2:-2]                                         # planes strictly between the two boundary faces
            assert not inner.any()
            assert not diff[0].any() and not diff[-1].any()            # ghosts along the normal untouched
            for plane, touched in ((diff[1], touched_lo), (diff[-2], touched_hi)):
                core = plane[1:-1, 1:-1]
                if touched:
                    assert np.ptp(core) <= 1e-15 * max(1.0, np.abs(core).max()) and core.flat[0] != 0.0
                    edge = plane.copy(); edge[1:-1, 1:-1] = 0
                    assert not edge.any()                              # ghosts of the other two directions untouched
                else:
                    assert not plane.any()


@pytest.mark.parametrize("name", sorted(CASES))
def test_solvability_matches_reference_kernels(name):
    spec = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    if "sol_eps_10" not in gold.files:
        pytest.skip("fixture predates the solvability vectors")
    for out_plane in (10, 1, 4):
        case = Case(tuple(spec["cells"]), bc=spec["bc"], nparts=spec.get("nparts", 0), radius=spec.get("radius", 1.0))
        _seed_star(case, MANIFEST["seed"] + 6)
        eps = case.o.solvability(out_plane)
        from cases import face_exchange_inputs
        fresh = {k: v[0] for k, v in face_exchange_inputs(case, 0, MANIFEST["seed"] + 6).items()}
        ref_eps = gold["sol_eps_%d" % out_plane]
        scale = sum(np.abs(v).sum() for v in fresh.values()) * 0.01
        assert np.abs(np.array(eps) - ref_eps).max() <= 1e-12 * scale
        for k, aid in (("u", ob.U_STAR), ("v", ob.V_STAR), ("w", ob.W_STAR)):
            if "sol_same_%s_%d" % (k, out_plane) in gold.files:          # the reference left this array alone
                assert np.array_equal(case.o.array(0, aid), fresh[k]), (k, out_plane)
                continue
            ref = gold["sol_%s_%d" % (k, out_plane)]
            assert np.abs(case.o.array(0, aid) - ref).max() <= 1e-12 * np.abs(ref).max(), (k, out_plane)
