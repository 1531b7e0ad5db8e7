"""GPU parity tests proper: the CUDA path (called through the C ABI) against the CPU oracle on
the same seeded inputs.  Tolerances: the reference path is FP64; north_star asks for pressure
within 1e-10 relative L2 and the residual history within +-1 iteration."""
import numpy as np
import pytest

from cases import Case, rel_l2
from oracle import binding as ob

pytestmark = pytest.mark.gpu

PHI_TOL = 1e-10          # relative L2 on the pressure correction (north_star)
HIST_TOL = 1e-7          # relative, per-iteration (r,z) (far tighter than +-1 iteration)
OP_TOL = 1e-13           # one application of the operator / rhs: FMA contraction only


def _product(case, **kw):
    from gpu_util import Product
    return Product(case, **kw)


@pytest.mark.parametrize("bc", ["cavity", "duct", "channel", "sedimentation", "periodic", "box"])
def test_rhs_matches_oracle(bc):
    case = Case((24, 20, 28), bc=bc)
    p = _product(case)
    p.each(lambda r, s, d: s.rhs(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"]))
    case.o.rhs(1.0, 1e-3)
    ref = case.o.array(0, ob.RHS_P)
    got = p.dev[0]["rhs_p"].cpu().numpy()
    assert np.abs(got - ref).max() <= OP_TOL * np.abs(ref).max()
    # ghosts are exactly zero (cudaMemset, cuda_solver.cu:122)
    g = got.copy(); g[1:-1, 1:-1, 1:-1] = 0
    assert not g.any()
    p.close()


@pytest.mark.parametrize("bc", ["cavity", "duct", "periodic"])
def test_spmv_noparts_matches_oracle(bc):
    case = Case((20, 24, 16), bc=bc)
    p = _product(case)
    p.set_coefficients()
    rng = np.random.default_rng(3)
    src = rng.standard_normal(case.o.array(0, ob.PB_Q).shape)
    case.o.array(0, ob.PB_Q)[...] = src
    case.o.spmv(ob.PB_Q, parts=False)
    ref = case.o.array(0, ob.APB_Q).copy()
    got = p.solvers[0].spmv(p.solvers[0].to_device(src)).cpu().numpy()
    assert np.abs(got - ref).max() <= OP_TOL * np.abs(ref).max()
    p.close()


def test_spmv_parts_matches_oracle():
    case = Case((32, 32, 32), bc="sedimentation", nparts=3, radius=2.0)
    assert (case.o.array(0, ob.PHASE) > -1).sum() > 50
    p = _product(case)
    p.set_coefficients(parts=True)
    rng = np.random.default_rng(4)
    src = rng.standard_normal(case.o.array(0, ob.PB_Q).shape)
    case.o.array(0, ob.PB_Q)[...] = src
    case.o.spmv(ob.PB_Q, parts=True)
    ref = case.o.array(0, ob.APB_Q).copy()
    got = p.solvers[0].spmv(p.solvers[0].to_device(src), use_phase=True).cpu().numpy()
    assert np.abs(got - ref).max() <= OP_TOL * np.abs(ref).max()
    p.close()


@pytest.mark.parametrize("blocks,bc", [((1, 1, 1), "periodic"), ((1, 1, 1), "cavity"), ((2, 1, 1), "channel"),
                                       ((2, 2, 1), "periodic"), ((1, 2, 2), "duct")])
def test_exchange_Gcc_matches_oracle(blocks, bc):
    """cuda_BC_test_periodic analogue (src/cuda_testing.cu:749-933): ghosts after the exchange."""
    case = Case((16, 12, 20), blocks=blocks, bc=bc)
    p = _product(case)
    rng = np.random.default_rng(5)
    for r in range(p.n):
        a = rng.standard_normal(case.o.array(r, ob.PHI).shape)
        case.o.array(r, ob.PHI)[...] = a
        p.dev[r]["phi"].copy_(p.solvers[r].to_device(a))
    case.o.exchange_Gcc(ob.PHI)
    p.each(lambda r, s, d: s.exchange_Gcc(d["phi"]))
    for r in range(p.n):
        assert np.array_equal(p.dev[r]["phi"].cpu().numpy(), case.o.array(r, ob.PHI)), "rank %d" % r
    p.close()


def _check_solve(case, p, parts=False, **kw):
    p.set_coefficients(parts=parts)
    res = p.solve(parts=parts, **kw)
    ores, ohist = case.solve_oracle(**{k: v for k, v in kw.items() if k in ("pp_residual", "pp_max_iter")})
    phi = p.gather("phi")
    ophi = case.o.gather_interior(ob.PHI)
    hist = p.solvers[0].history()
    for r in res:
        assert r.status == {0: "converged", 1: "tiny_rhs", 2: "max_iter", 3: "nan"}[ores.status]
        assert abs(r.niter - ores.niter) <= 1, (r.niter, ores.niter)
    n = min(len(hist), len(ohist))
    assert n >= min(res[0].niter, ores.niter)
    assert np.max(np.abs(hist[:n] - ohist[:n]) / ohist[:n]) < HIST_TOL
    assert abs(res[0].sp_rhs - ores.sp_rhs) <= 1e-12 * ores.sp_rhs
    if res[0].niter == ores.niter:
        assert rel_l2(phi, ophi) < PHI_TOL
        assert abs(res[0].resid - ores.resid) <= 1e-6 * ores.resid
    return res, ores, phi, ophi


@pytest.mark.parametrize("bc", ["cavity", "duct", "channel", "sedimentation", "periodic", "box"])
def test_solve_noparts_all_bc_sets(bc):
    case = Case((32, 32, 32), bc=bc)
    p = _product(case)
    res, ores, _, _ = _check_solve(case, p)
    assert res[0].niter > 50          # crosses the q % 50 true-residual refresh
    p.close()


@pytest.mark.parametrize("ty", [1, 2, 3, 4, 5, 6, 7, 8])
def test_solve_every_tile_height(ty):
    """the tile height of the two iteration kernels is a run-time argument picked by the planner (fills the SM slots)"""
    case = Case((40, 24, 36), bc="duct")      # ragged: not a multiple of any tile (general form of the plane loop)
    p = _product(case, options={"ty": ty})
    assert p.solvers[0].info("search_ty") == ty
    _check_solve(case, p)
    p.close()


@pytest.mark.parametrize("ty,kc", [(0, 0), (7, 9), (5, 11), (8, 1000)])
@pytest.mark.parametrize("cells,blocks,parts", [((128, 28, 44), (1, 1, 1), False), ((256, 20, 24), (2, 1, 2), False),
                                                ((128, 36, 28), (1, 1, 1), True), ((128, 36, 28), (1, 2, 1), True),
                                                ((36, 28, 44), (2, 1, 2), False), ((36, 28, 44), (1, 2, 1), True)])
def test_full_and_ragged_tiles(ty, kc, cells, blocks, parts):
    """in a multiple of 128: the predicate-free XFULL form of both plane loops (with its all-ones-mask fast path next to
    walls and particles, where it must fall back per warp); other sizes: the general form.  Decomposed and with particles."""
    kw = dict(nparts=3, radius=2.5) if parts else {}
    ext = tuple(v for n in cells for v in (0., n / 3.))          # dx = 1/3 in every direction: a sphere is 7.5 cells in radius
    case = Case(cells, blocks=blocks, bc="sedimentation" if parts else "channel", extent=ext, **kw)
    p = _product(case, options={"ty": ty, "kc": kc})
    _check_solve(case, p, parts=parts)
    p.close()


@pytest.mark.parametrize("opts", [{"kc": 1}, {"kc": 3, "ty": 4}, {"kc": 1000}, {"pdl": 0}, {"pdl": 1}, {"rhs_tiled": 0}])
def test_zchunk_plans(opts):
    case = Case((24, 20, 52), bc="cavity")      # > 100 iterations: two true-residual refreshes
    p = _product(case, options=opts)
    _check_solve(case, p)
    p.close()


def test_plan_of_the_benchmarked_block_shapes():
    """the planner's tile / z-chunk choice for the benchmarked block shapes: 8-row tiles, ONE wave of resident
    CTAs (2 per SM) claiming the work items; uniform ~24-plane chunks when there are many items, chunks of DECREASING length
    (never below chunk_min) for the small blocks of an 8-GPU run"""
    import bbpcg
    from bbpcg.grid import BC_SETS
    for cells, many in (((256, 256, 256), False), ((512, 512, 64), False), ((512, 512, 512), True)):
        dec = bbpcg.Decomposition.uniform((0., 12., 0., 12., 0., 12.), cells, (1, 1, 1), BC_SETS["duct"])
        s = bbpcg.PoissonSolver(dec, 0)
        slots = 2 * s.info("sm_count")
        grid, items, ty, kc, nbz = s.info("search_grid"), s.info("search_items"), s.info("search_ty"), s.info("search_kc"), s.info("search_nbz")
        assert ty == 8 and items == (cells[0] // 128) * -(-cells[1] // ty) * nbz and grid == min(items, slots)
        if many:
            assert items >= 7 * slots and kc == 24 and nbz == -(-cells[2] // kc)
        else:
            assert kc >= -(-cells[2] // nbz) and (nbz >= 2 or kc == cells[2])     # guided: the first chunk is the longest
        s.set_option("ty", 7); s.set_option("kc", 16)
        assert s.info("search_ty") == 7 and s.info("search_kc") == 16 and s.info("search_nbz") == -(-cells[2] // 16)
        s.close()


@pytest.mark.parametrize("cells", [(33, 17, 9), (7, 5, 3), (130, 9, 18), (16, 16, 130)])
def test_solve_ragged_and_odd_sizes(cells):
    case = Case(cells, bc="channel")
    p = _product(case, options={"kc": 7})
    _check_solve(case, p)
    p.close()


def test_solve_lid_driven_cavity_96():
    """BASELINE configs[0]: examples/lid-driven-cavity, 96^3, 1 rank."""
    case = Case((96, 96, 96), bc="cavity", omp=True)
    p = _product(case)
    res, ores, phi, ophi = _check_solve(case, p)
    assert res[0].niter == ores.niter
    p.close()


def test_solve_with_particles():
    case = Case((40, 40, 40), bc="sedimentation", nparts=4, radius=2.5)
    p = _product(case)
    res, ores, phi, ophi = _check_solve(case, p, parts=True)
    # phi stays 0 inside the particles (rhs = r = p = 0 there for the whole solve)
    solid = case.o.gather_interior(ob.PHASE) > -1
    assert solid.sum() > 100 and np.all(phi[solid] == 0.0)
    p.close()


@pytest.mark.parametrize("blocks,bc", [((2, 1, 1), "channel"), ((1, 1, 2), "channel"), ((2, 2, 1), "cavity"),
                                       ((2, 2, 2), "sedimentation"), ((2, 2, 2), "periodic"), ((1, 2, 1), "box")])
def test_solve_decomposed_matches_single_block_oracle(blocks, bc):
    """The discrete solution is decomposition independent: N ranks vs the 1-block oracle."""
    case = Case((32, 32, 32), blocks=blocks, bc=bc)
    single = Case((32, 32, 32), bc=bc)
    p = _product(case)
    p.set_coefficients()
    res = p.solve()
    ores, ohist = single.solve_oracle()
    phi, ophi = p.gather("phi"), single.o.gather_interior(ob.PHI)
    assert all(r.niter == res[0].niter and r.status == "converged" for r in res)
    assert abs(res[0].niter - ores.niter) <= 1
    hist = p.solvers[0].history()
    n = min(len(hist), len(ohist))
    assert np.max(np.abs(hist[:n] - ohist[:n]) / ohist[:n]) < HIST_TOL
    if res[0].niter == ores.niter:
        assert rel_l2(phi, ophi) < PHI_TOL
    # every rank made identical decisions from the rank-ordered all-reduce
    for s in p.solvers[1:]:
        assert np.array_equal(s.history(), hist)
    p.close()


def test_decomposed_with_particles():
    case = Case((32, 32, 32), blocks=(2, 2, 1), bc="sedimentation", nparts=3, radius=2.5)
    p = _product(case)
    _check_solve(case, p, parts=True)
    p.close()


def test_run_to_run_bit_identical():
    case = Case((32, 32, 32), bc="cavity")
    p = _product(case)
    p.set_coefficients()
    p.solve(); h1 = p.solvers[0].history(); phi1 = p.gather("phi")
    p.solve(); h2 = p.solvers[0].history(); phi2 = p.gather("phi")
    assert np.array_equal(h1, h2) and np.array_equal(phi1, phi2)
    p.close()


def test_tiny_rhs_shortcut_and_maxiter_and_fixed():
    case = Case((16, 16, 16), bc="periodic")
    p = _product(case)
    p.set_coefficients()
    d = p.dev[0]
    z = {k: d[k] * 0 for k in ("u_star", "v_star", "w_star")}
    s = p.solvers[0]
    r = s.PP_cg_noparts(z["u_star"], z["v_star"], z["w_star"], d["rhs_p"], d["phi"])
    assert r.status == "tiny_rhs" and r.niter == 0 and not d["phi"].any()          # cuda_solver.cu:641-651
    r = s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], pp_max_iter=5)
    ores, _ = case.solve_oracle(pp_max_iter=5)
    assert r.status == "max_iter" and r.niter == 6 == ores.niter                   # loop bound q <= pp_max_iter, :654
    r = s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"], fixed_iters=23)
    assert r.niter == 23 and len(s.history()) == 24
    p.close()


def test_solve_host_buffers_end_to_end():
    import torch
    case = Case((24, 24, 24), bc="duct")
    p = _product(case)
    p.set_coefficients()
    s, inp = p.solvers[0], case.inputs(0)
    hu, hv, hw = [torch.from_numpy(inp[k].copy()).pin_memory() for k in ("u_star", "v_star", "w_star")]
    hphi = torch.zeros(tuple(case.o.array(0, ob.PHI).shape), dtype=torch.float64).pin_memory()
    r = s.solve_host(hu, hv, hw, hphi)
    ores, _ = case.solve_oracle()
    assert r.niter == ores.niter
    assert rel_l2(hphi.numpy()[1:-1, 1:-1, 1:-1], case.o.gather_interior(ob.PHI)) < PHI_TOL
    p.close()


@pytest.mark.parametrize("n", [128, 256])
def test_size_independent_properties(n):
    """Properties that need no oracle run: true residual of the returned phi, linearity in rhs.  256^3 duct is BASELINE.json
    configs[1] (examples/pressure-driven-duct, particle-free, one B200)."""
    import torch
    case = Case((n, n, n), bc="duct", omp=True)
    p = _product(case)
    p.set_coefficients()
    s, d = p.solvers[0], p.dev[0]
    r1 = s.PP_cg_noparts(d["u_star"], d["v_star"], d["w_star"], d["rhs_p"], d["phi"])
    assert r1.status == "converged"
    phi1 = d["phi"].clone()
    b = d["rhs_p"][1:-1, 1:-1, 1:-1].clone()
    s.exchange_Gcc(d["phi"])
    Aphi = s.spmv(d["phi"])
    true_res = float(torch.linalg.norm(b - Aphi) / torch.linalg.norm(b))
    dx = 12.0 / n
    assert true_res < 2.0 * (6 ** 0.5 / dx) * 1e-6      # the reported 1e-6 is the preconditioned norm, ~dx/sqrt(6) smaller
    # linearity: solving for 2*u* gives 2*phi with the same iteration count
    r2 = s.PP_cg_noparts(2 * d["u_star"], 2 * d["v_star"], 2 * d["w_star"], d["rhs_p"], d["phi"])
    assert r2.niter == r1.niter
    assert float(torch.linalg.norm(d["phi"][1:-1, 1:-1, 1:-1] - 2 * phi1[1:-1, 1:-1, 1:-1]) /
                 torch.linalg.norm(phi1[1:-1, 1:-1, 1:-1])) < 1e-12
    p.close()
