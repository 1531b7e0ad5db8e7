"""GPU parity of the solve epilogue (bbpcg_dom_BC_p / bbpcg_epilogue through the C ABI):
exchange(phi) + cuda_dom_BC_p(phi) + cuda_project + cuda_update_p, src/bluebottle.c:233-250.

Three-way where the reference library is present: the reference's own kernels (O1, oracle/_ref/libbbref.so) vs the
CPU oracle (O2, oracle/pcg_ref.c) vs the product.  Tolerances: ghost fills are copies -> bit-exact; u, v, w differ
from O1 by nothing but possible FMA-contraction choices (1e-14 of the field's max; O2 is compiled without
contraction); p by the summation order of the mean (1e-12)."""
import ctypes as C

import numpy as np
import pytest

from cases import Case, load_ref, ref_epilogue
from oracle import binding as ob

pytestmark = pytest.mark.gpu

VEL_TOL = 1e-14
P_TOL = 1e-12
INNER = (slice(1, -1),) * 3


def _product(case, options=None):
    from gpu_util import Product
    p = Product(case, options)
    for r, s in enumerate(p.solvers):
        d = p.dev[r]
        ein = case.epilogue_inputs(r)
        d["phi"] = s.to_device(ein["phi"])
        d["p0"] = s.to_device(ein["p0"])
        d["u"], d["v"], d["w"], d["p"] = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
    return p


def _fused(p, **kw):
    return p.each(lambda r, s, d: s.epilogue(d["phi"], d["u_star"], d["v_star"], d["w_star"], d["flag_u"], d["flag_v"], d["flag_w"],
                                             d["u"], d["v"], d["w"], d["p0"], d["phase"], d["p"], **kw))


def _close(a, b, tol):
    return np.abs(a - b).max() <= tol * np.abs(b).max()


def _check_against_oracle(case, p):
    for r in range(p.n):
        dev = p.dev[r]
        assert np.array_equal(dev["phi"].cpu().numpy(), case.o.array(r, ob.PHI))       # ghost faces filled, nothing else touched
        for key, aid in (("u", ob.U), ("v", ob.V), ("w", ob.W)):
            assert _close(dev[key].cpu().numpy()[INNER], case.o.array(r, aid)[INNER], VEL_TOL), (r, key)
        assert _close(dev["p"].cpu().numpy()[INNER], case.o.array(r, ob.P)[INNER], P_TOL), r


EPI_FORMS = {0: {"epilogue_tiled": 0},      # the plane-marching pair k_epi_uwp + k_epi_v (default)
             1: {"epilogue_tiled": 1}}      # the 16^3-brick kernel k_epilogue


@pytest.mark.parametrize("tiled", [0, 1])
@pytest.mark.parametrize("cells,bc,nparts", [((32, 28, 36), "cavity", 0), ((32, 28, 36), "duct", 0), ((32, 28, 36), "periodic", 0),
                                             ((33, 17, 9), "box", 0), ((16, 16, 16), "channel", 0), ((40, 40, 40), "sedimentation", 3),
                                             ((70, 33, 45), "box", 0), ((64, 96, 65), "duct", 0)])
def test_epilogue_three_way(cells, bc, nparts, tiled):
    """both forms of the epilogue (EPI_FORMS) against the oracle and the reference's kernels"""
    case = Case(cells, bc=bc, nparts=nparts, radius=2.5)
    case.seed_epilogue(23)
    ein = case.epilogue_inputs(0)
    p = _product(case, options=dict(EPI_FORMS[tiled], epi_chunk=0 if cells[0] != 64 else 24))
    _fused(p)
    case.o.epilogue(1.0, 1e-3)
    _check_against_oracle(case, p)
    lib = load_ref()
    if lib is not None:                                   # the reference's own kernels
        dom, DOM = case.o.dom(0), case.o.DOM
        assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
        P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        k = {n: np.ascontiguousarray(v) for n, v in case.inputs(0).items()}
        assert lib.bbref_set_inputs(P(k["flag_u"]), P(k["flag_v"]), P(k["flag_w"]), P(k["phase"]), P(k["phase_shell"]),
                                    P(k["u_star"]), P(k["v_star"]), P(k["w_star"]), nparts) == 0
        ref = ref_epilogue(lib, case, ein["phi"], ein["p0"])
        dev = p.dev[0]
        assert np.array_equal(dev["phi"].cpu().numpy(), ref["phi"])
        for key in ("u", "v", "w"):
            assert _close(dev[key].cpu().numpy()[INNER], ref[key][INNER], VEL_TOL), key
            assert _close(case.o.array(0, getattr(ob, key.upper()))[INNER], ref[key][INNER], VEL_TOL), key
        assert _close(dev["p"].cpu().numpy()[INNER], ref["p"][INNER], P_TOL)
        assert _close(case.o.array(0, ob.P)[INNER], ref["p"][INNER], P_TOL)
    p.close()


@pytest.mark.parametrize("blocks,bc", [((2, 1, 1), "duct"), ((1, 2, 2), "cavity"), ((2, 2, 2), "periodic"), ((1, 1, 3), "sedimentation"),
                                       ((2, 2, 1), "channel")])
def test_epilogue_decomposed(blocks, bc):
    """several ranks (one process, one GPU; same kernels, peer stores and in-kernel all-reduce as a multi-GPU run)"""
    case = Case((36, 34, 30), blocks=blocks, bc=bc)
    case.seed_epilogue(31)
    p = _product(case)
    _fused(p)
    case.o.epilogue(1.0, 1e-3)
    _check_against_oracle(case, p)
    # every rank subtracted the same global mean: the assembled field has zero mean
    full = p.gather("p")
    assert abs(full.mean()) <= 1e-13 * np.abs(full).max()
    p.close()


def test_epilogue_streaming_pair_equals_the_brick_kernel():
    """same expressions, different traversal: u, v, w bit-identical; p differs only by the summation order of its mean"""
    case = Case((70, 40, 50), bc="cavity")
    case.seed_epilogue(29)
    outs = []
    for tiled in (0, 1):
        p = _product(case, options=EPI_FORMS[tiled])
        _fused(p)
        outs.append({k: p.dev[0][k].cpu().numpy().copy() for k in ("u", "v", "w", "p")})
        p.close()
    for o in outs[1:]:
        for k in ("u", "v", "w"):
            assert np.array_equal(outs[0][k][INNER], o[k][INNER]), k
        assert np.abs(outs[0]["p"][INNER] - o["p"][INNER]).max() <= 1e-13 * np.abs(o["p"][INNER]).max()


def test_epilogue_halves_equal_the_fused_call():
    """mpi_cuda_exchange_Gcc + cuda_dom_BC_p + cuda_project + cuda_update_p called one by one (the drop-in order,
    src/bluebottle.c:233-250) give bit-identical arrays to the single fused call"""
    case = Case((40, 24, 20), blocks=(2, 1, 1), bc="cavity")
    case.seed_epilogue(41)
    a, b = _product(case), _product(case)
    _fused(a)
    b.each(lambda r, s, d: s.exchange_Gcc(d["phi"]))
    b.each(lambda r, s, d: s.dom_BC_p(d["phi"]))
    b.each(lambda r, s, d: s.project(d["u_star"], d["v_star"], d["w_star"], d["phi"], d["flag_u"], d["flag_v"], d["flag_w"],
                                     d["u"], d["v"], d["w"]))
    b.each(lambda r, s, d: s.update_p(d["p0"], d["phi"], d["phase"], d["p"]))
    for r in range(a.n):
        for key in ("phi", "u", "v", "w", "p"):
            assert np.array_equal(a.dev[r][key].cpu().numpy(), b.dev[r][key].cpu().numpy()), (r, key)
    # run to run
    _fused(b, phi_ghosts_valid=True)
    for r in range(a.n):
        assert np.array_equal(a.dev[r]["p"].cpu().numpy(), b.dev[r]["p"].cpu().numpy())
    a.close(); b.close()


def test_epilogue_argument_errors():
    import bbpcg
    case = Case((8, 8, 8), bc="duct")
    case.seed_epilogue(1)
    p = _product(case)
    s, d = p.solvers[0], p.dev[0]
    with pytest.raises(RuntimeError, match="nothing to do"):
        s.epilogue(d["phi"])
    with pytest.raises(RuntimeError, match="cuda_project needs"):
        s.epilogue(d["phi"], u=d["u"])
    with pytest.raises(RuntimeError, match="cuda_update_p needs"):
        s.epilogue(d["phi"], p=d["p"])
    p.close()


def test_solve_then_epilogue_128_properties():
    """BASELINE-sized behaviour through size-independent properties: after the solve the projected velocity is
    discretely divergence free to the solve tolerance, and the updated pressure has zero mean"""
    import torch
    case = Case((128, 128, 128), bc="duct", omp=True)
    case.seed_epilogue(3, phi=False)
    p = _product(case)
    p.set_coefficients()
    res = p.solve(pp_residual=1e-9)[0]
    assert res.status == "converged"
    ms = _fused(p)[0]
    assert ms > 0
    d = case.o.dom(0)
    dev = p.dev[0]

    def div(u, v, w):
        u, v, w = u[INNER], v[INNER], w[INNER]
        return ((u[1:] - u[:-1]).permute(1, 2, 0) / d.dx + (v[1:] - v[:-1]).permute(2, 0, 1) / d.dy + (w[1:] - w[:-1]) / d.dz)
    before = div(dev["u_star"], dev["v_star"], dev["w_star"])
    after = div(dev["u"], dev["v"], dev["w"])
    assert float(torch.linalg.norm(after)) < 1e-5 * float(torch.linalg.norm(before))
    pin = dev["p"][INNER]
    assert abs(float(pin.mean())) <= 1e-12 * float(pin.abs().max())
    # p = p0 + phi - mean in every (fluid) cell
    expect = (dev["p0"] + dev["phi"])[INNER]
    expect = expect - expect.mean()
    assert float((pin - expect).abs().max()) <= 1e-11 * float(expect.abs().max())
    p.close()


# ---- face-grid halo exchanges: bbpcg_exchange = mpi_cuda_exchange_Gfx / _Gfy / _Gfz (src/mpi_comm.c:317-405) -----
FACE = {"u": ("Gfx", ob.U), "v": ("Gfy", ob.V), "w": ("Gfz", ob.W)}


@pytest.mark.parametrize("blocks,bc", [((1, 1, 1), "periodic"), ((1, 1, 1), "duct"), ((2, 1, 1), "periodic"), ((1, 2, 2), "channel"),
                                       ((2, 2, 2), "periodic"), ((3, 1, 2), "sedimentation")])
def test_face_exchange_matches_oracle(blocks, bc):
    """pure copies: bit-exact against the oracle's multi-block exchange, ghost edges/corners and wall ghosts untouched"""
    from cases import face_exchange_inputs
    from gpu_util import Product
    case = Case((24, 14, 18), blocks=blocks, bc=bc)
    p = Product(case)
    for key, (grid, aid) in FACE.items():
        for r in range(p.n):
            arr = face_exchange_inputs(case, r, 61)[key][0]
            case.o.array(r, aid)[...] = arr
            p.dev[r]["x" + key] = p.solvers[r].to_device(arr)
        case.o.exchange(aid)
        p.each(lambda r, s, d: s.exchange(d["x" + key], grid))
        for r in range(p.n):
            assert np.array_equal(p.dev[r]["x" + key].cpu().numpy(), case.o.array(r, aid)), (key, r)
    # the Gcc entry point is the grid code 0 of the same kernels
    for r in range(p.n):
        arr = np.random.default_rng(5 + r).standard_normal(case.o.array(r, ob.PHI).shape)
        case.o.array(r, ob.PHI)[...] = arr
        p.dev[r]["xc"] = p.solvers[r].to_device(arr)
    case.o.exchange_Gcc(ob.PHI)
    p.each(lambda r, s, d: s.exchange(d["xc"], "Gcc"))
    for r in range(p.n):
        assert np.array_equal(p.dev[r]["xc"].cpu().numpy(), case.o.array(r, ob.PHI))
    p.close()


@pytest.mark.parametrize("bc", ["periodic", "channel"])
def test_face_exchange_three_way(bc):
    """the reference's own pack / unpack kernels (O1) vs the oracle vs the product, single block (periodic self-wrap)"""
    from cases import face_exchange_inputs
    from gpu_util import Product
    lib = load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libbbref.so not built")
    case = Case((20, 12, 16), bc=bc)
    dom, DOM = case.o.dom(0), case.o.DOM
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    p = Product(case)
    s = p.solvers[0]
    for key, (arr, code) in face_exchange_inputs(case, 0, 67).items():
        grid, aid = FACE[key]
        ref = np.ascontiguousarray(arr).copy()
        assert lib.bbref_exchange_face(ref.ctypes.data_as(C.c_void_p), code) == 0
        case.o.array(0, aid)[...] = arr
        case.o.exchange(aid)
        mine = s.to_device(arr)
        s.exchange(mine, grid)
        assert np.array_equal(case.o.array(0, aid), ref), key
        assert np.array_equal(mine.cpu().numpy(), ref), key
    p.close()


def test_replay_from_a_restart_file(tmp_path):
    """a Bluebottle restart file (out_restart, src/domain.c:3005-3092) replayed through the library:
    flags, phase and u* come from the file; the solve and the epilogue equal the run fed from the arrays directly"""
    import bbpcg
    from gpu_util import Product
    from test_restart import write_restart
    case = Case((24, 20, 28), bc="duct")
    case.seed_epilogue(13, phi=False)
    d0 = case.o.dom(0)
    inp = case.inputs(0)
    fields = dict(u=np.zeros_like(inp["u_star"]), v=np.zeros_like(inp["v_star"]), w=np.zeros_like(inp["w_star"]),
                  u_star=inp["u_star"], v_star=inp["v_star"], w_star=inp["w_star"], p=case.o.array(0, ob.P),
                  phi=case.o.array(0, ob.PHI), p0=case.o.array(0, ob.P0), phase=inp["phase"], phase_shell=inp["phase_shell"],
                  flag_u=inp["flag_u"], flag_v=inp["flag_v"], flag_w=inp["flag_w"])
    path = bbpcg.restart_path(str(tmp_path), 0, 1)
    write_restart(path, d0, fields, dt=1e-3)
    rst = bbpcg.read_restart(path, d0)
    p = Product(case)
    s = p.solvers[0]
    dev = {k: s.to_device(rst[k]) for k in ("u_star", "v_star", "w_star", "flag_u", "flag_v", "flag_w", "phase", "p0")}
    rhs, phi = s.empty("Gcc"), s.empty("Gcc")
    un, vn, wn, pn = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
    s.init_jacobi_preconditioner(dev["flag_u"], dev["flag_v"], dev["flag_w"])
    res = s.PP_cg_noparts(dev["u_star"], dev["v_star"], dev["w_star"], rhs, phi, dt=rst["dt"])
    s.epilogue(phi, dev["u_star"], dev["v_star"], dev["w_star"], dev["flag_u"], dev["flag_v"], dev["flag_w"], un, vn, wn,
               dev["p0"], dev["phase"], pn, dt=rst["dt"])
    ores, _ = case.solve_oracle()
    case.o.epilogue(1.0, 1e-3)
    assert res.status == "converged" and res.niter == ores.niter
    from cases import rel_l2
    assert rel_l2(phi.cpu().numpy()[INNER], case.o.array(0, ob.PHI)[INNER]) < 1e-10
    assert np.abs(un.cpu().numpy()[INNER] - case.o.array(0, ob.U)[INNER]).max() < 1e-9
    assert np.abs(pn.cpu().numpy()[INNER] - case.o.array(0, ob.P)[INNER]).max() < 1e-9 * np.abs(case.o.array(0, ob.P)).max()
    p.close()


# ---- cuda_solvability (src/cuda_bluebottle.cu:2313-2492): bbpcg_solvability -----------------------------------------
@pytest.mark.parametrize("blocks,out_plane", [((1, 1, 1), "HOMOGENEOUS"), ((1, 1, 1), "EAST"), ((1, 1, 1), "BOTTOM"), ((2, 1, 2), "HOMOGENEOUS"),
                                              ((2, 2, 1), "NORTH"), ((1, 3, 1), "WEST"), ((2, 2, 2), "TOP")])
def test_solvability_matches_oracle_and_reference(blocks, out_plane):
    """sums differ from the oracle's / Thrust's only in summation order (1e-12 of the summed magnitude); the planes that
    must not change are bit-identical; afterwards the net boundary flux is zero"""
    from bbpcg.lib import OUT_PLANE
    from cases import face_exchange_inputs
    from gpu_util import Product
    case = Case((24, 18, 16), blocks=blocks, bc="box")
    p = Product(case)
    scale = 0.0
    for r in range(p.n):
        fx = face_exchange_inputs(case, r, 71)
        for key, aid, name in (("u", ob.U_STAR, "u_star"), ("v", ob.V_STAR, "v_star"), ("w", ob.W_STAR, "w_star")):
            case.o.array(r, aid)[...] = fx[key][0]
            p.dev[r][name] = p.solvers[r].to_device(fx[key][0])
            scale += np.abs(fx[key][0]).sum() * 0.01
    oeps = case.o.solvability(OUT_PLANE[out_plane])
    eps = p.each(lambda r, s, d: s.solvability(d["u_star"], d["v_star"], d["w_star"], out_plane))
    for r in range(p.n):
        assert eps[r] == eps[0]                                      # every rank holds the same bits
        assert np.abs(np.array(eps[r]) - np.array(oeps)).max() <= 1e-12 * scale
        for aid, name in ((ob.U_STAR, "u_star"), (ob.V_STAR, "v_star"), (ob.W_STAR, "w_star")):
            mine, ref = p.dev[r][name].cpu().numpy(), case.o.array(r, aid)
            assert np.abs(mine - ref).max() <= 1e-12 * np.abs(ref).max(), (r, name)
            same = ref == face_exchange_inputs(case, r, 71)[name[0]][0]
            assert np.array_equal(mine[same], ref[same])            # untouched entries are untouched
    if blocks == (1, 1, 1):
        lib = load_ref()
        if lib is not None:
            dom, DOM = case.o.dom(0), case.o.DOM
            assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
            fx = face_exchange_inputs(case, 0, 71)
            arrs = {k: np.ascontiguousarray(v[0]).copy() for k, v in fx.items()}
            reps = (C.c_double * 3)()
            P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
            assert lib.bbref_solvability(P(arrs["u"]), P(arrs["v"]), P(arrs["w"]), OUT_PLANE[out_plane], reps) == 0
            assert np.abs(np.array(list(reps)) - np.array(eps[0])).max() <= 1e-12 * scale
            for k, name in (("u", "u_star"), ("v", "v_star"), ("w", "w_star")):
                assert np.abs(p.dev[0][name].cpu().numpy() - arrs[k]).max() <= 1e-12 * np.abs(arrs[k]).max(), k
    p.close()
