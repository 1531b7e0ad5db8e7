"""Parity at the sizes that are BENCHMARKED (BASELINE.json configs), against the reference's OWN kernels (O1,
oracle/_ref/libbbref.so = /root/reference/src hot-path TUs unmodified + a single-rank shim): the discrete solution does not
depend on the decomposition, so the 1-rank reference solve of the global grid is also the oracle of the decomposed runs
(SURVEY.md 8c).  Inputs are built on the GPU by the same deterministic generators bench.py uses.  Tolerances (north_star):
phi within 1e-10 relative L2, iteration count within +-1 (observed: equal)."""
import ctypes as C

import numpy as np
import pytest

from cases import load_ref
from oracle import binding as ob

pytestmark = pytest.mark.gpu

PHI_TOL = 1e-10


def _reference_solve(extent, cells, bcname, inputs, nparts=0):
    """the reference's cuda_PP_init_jacobi_preconditioner + cuda_PP_cg[_noparts] on ONE block; returns niter, resid, phi (GPU tensor)"""
    import torch
    from bbpcg.grid import BC_SETS, grid_shape
    lib = load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference at build time)")
    DOM, dom, _ = ob.single_block_domain(extent, cells, BC_SETS[bcname])
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    i = inputs
    torch.cuda.synchronize()
    assert lib.bbref_set_inputs_dev(P(i["flag_u"]), P(i["flag_v"]), P(i["flag_w"]), P(i["phase"]), P(i["phase_shell"]),
                                    P(i["u_star"]), P(i["v_star"]), P(i["w_star"]), nparts) == 0
    niter, resid, ms = C.c_int(), C.c_double(), C.c_float()
    assert lib.bbref_solve(1.0, 1e-3, 1e-6, 2000, 1 if nparts else 0, C.byref(niter), C.byref(resid), C.byref(ms)) == 0
    phi = np.zeros(grid_shape(dom, "Gcc"))
    assert lib.bbref_get(0, phi.ctypes.data_as(C.c_void_p)) == 0
    return niter.value, resid.value, torch.from_numpy(phi[1:-1, 1:-1, 1:-1].copy()).cuda()


def _rel_l2(a, b):
    return float(((a - b) ** 2).sum().sqrt() / (b ** 2).sum().sqrt())


@pytest.mark.parametrize("cells,bc,blocks", [((256, 256, 256), "duct", (1, 1, 1)),          # BASELINE configs[1]
                                             ((512, 256, 256), "channel", (1, 1, 1)),        # configs[2], 1 GPU
                                             ((512, 256, 256), "channel", (2, 1, 1)),        # x split: element-strided E/W faces
                                             ((512, 256, 256), "channel", (1, 1, 2)),        # z split: contiguous T/B faces
                                             ((256, 256, 256), "duct", (2, 2, 2))])          # the 8-rank shape of the scaling runs
def test_benchmarked_configs_match_the_reference_kernels(cells, bc, blocks):
    from gpu_util import TorchProduct
    extent = (0., 12., 0., 12. * cells[1] / cells[0], 0., 12. * cells[2] / cells[0])
    one = TorchProduct(extent, cells, (1, 1, 1), bc)            # the 1-block inputs feed the reference
    rn, rres, rphi = _reference_solve(extent, cells, bc, one.dev[0])
    p = one if blocks == (1, 1, 1) else TorchProduct(extent, cells, blocks, bc)
    p.set_coefficients()
    res = p.solve()
    for r in res:
        assert r.status == "converged" and abs(r.niter - rn) <= 1, (r, rn)
    assert res[0].niter == rn                                   # observed: equal
    assert abs(res[0].resid - rres) <= 1e-6 * rres
    assert _rel_l2(p.gather("phi"), rphi) < PHI_TOL
    p.close()
    if p is not one:
        one.close()


def test_particles_at_scale_match_the_reference_kernels():
    """cuda_PP_cg with >= 100 spheres on >= 128^3 (the scaled-down shape of BASELINE configs[3]: sedimentation set, 8 cells
    per radius): product (1 block and 2 x 1 x 2 blocks) against the reference's own kernels"""
    from gpu_util import TorchProduct
    cells, bc, nparts = (128, 128, 160), "sedimentation", 110
    extent = (0., 16., 0., 16., 0., 20.)                        # dx = 1/8, radius 1 = 8 cells
    one = TorchProduct(extent, cells, (1, 1, 1), bc, nparts=nparts)
    solid = int((one.dev[0]["phase"][1:-1, 1:-1, 1:-1] > -1).sum())
    assert solid > 100 * 1500                                   # ~2145 cells per sphere
    rn, rres, rphi = _reference_solve(extent, cells, bc, one.dev[0], nparts=nparts)
    for blocks in ((1, 1, 1), (2, 1, 2)):
        p = one if blocks == (1, 1, 1) else TorchProduct(extent, cells, blocks, bc, nparts=nparts)
        p.set_coefficients()
        res = p.solve()
        assert all(r.status == "converged" and abs(r.niter - rn) <= 1 for r in res), (res, rn)
        if res[0].niter == rn:
            assert _rel_l2(p.gather("phi"), rphi) < PHI_TOL
        phi = p.gather("phi")
        assert float(phi[one.dev[0]["phase"][1:-1, 1:-1, 1:-1] > -1].abs().max()) == 0.0      # solid cells stay exactly 0
        p.close()
    one.close()


@pytest.mark.parametrize("blocks", [(1, 1, 1), (1, 2, 1)])
def test_nan_exit(blocks):
    """src/cuda_solver.cu:245-251 (a16): a NaN in u* makes (b,b) and (r,z) NaN; the tiny-rhs test (NaN < tol is false) lets it
    through, the first iteration's `isnan(sp_rq1)` stops the solve: status nan after exactly 1 iteration, on every rank,
    and the oracle agrees"""
    from cases import Case
    from gpu_util import Product
    case = Case((24, 20, 28), blocks=blocks, bc="duct")
    case.o.array(0, ob.U_STAR)[5, 7, 3] = np.nan
    ores, _ = case.solve_oracle()
    assert ores.status == 3 and ores.niter == 1
    p = Product(case)
    p.set_coefficients()
    res = p.solve()
    for r in res:
        assert r.status == "nan" and r.niter == 1 and np.isnan(r.resid), r
    # the solver object stays usable: the same solve without the NaN converges
    case2 = Case((24, 20, 28), blocks=blocks, bc="duct")
    for r in range(p.n):
        p.dev[r]["u_star"].copy_(p.solvers[r].to_device(case2.o.array(r, ob.U_STAR)))
    res = p.solve()
    ores2, _ = case2.solve_oracle()
    assert all(r.status == "converged" and r.niter == ores2.niter for r in res)
    p.close()
