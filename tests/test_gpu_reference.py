"""Three-way parity on the GPU box: the reference's OWN kernels (O1: oracle/_ref/libbbref.so, the
unmodified /root/reference/src hot-path TUs) vs the CPU oracle (O2: oracle/pcg_ref.c) vs the CUDA
product (through the C ABI), on the same seeded inputs.  This is what pins the oracle: the
reference ships no golden vector for the Poisson solver."""
import ctypes as C

import numpy as np
import pytest

from cases import Case, load_ref, rel_l2
from oracle import binding as ob

pytestmark = pytest.mark.gpu

PHI_TOL = 1e-10      # north_star: pressure within 1e-10 relative L2
OP_TOL = 1e-13       # one operator application: FMA contraction / association only


def _P(a):
    return a.ctypes.data_as(C.c_void_p)


class Ref:
    """One reference state (globals inside libbbref.so: one live grid at a time)."""

    def __init__(self, case, nparts=0):
        self.lib = load_ref()
        if self.lib is None:
            pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference at build time)")
        self.case = case
        self.dom, self.DOM = case.o.dom(0), case.o.DOM
        assert self.lib.bbref_init(C.byref(self.dom), C.byref(self.DOM)) == 0
        self.inp = {k: np.ascontiguousarray(v) for k, v in case.inputs(0).items()}
        i = self.inp
        assert self.lib.bbref_set_inputs(_P(i["flag_u"]), _P(i["flag_v"]), _P(i["flag_w"]), _P(i["phase"]), _P(i["phase_shell"]),
                                         _P(i["u_star"]), _P(i["v_star"]), _P(i["w_star"]), nparts) == 0
        g = self.dom.Gcc
        self.s3b = (g.get("knb"), g.get("jnb"), g.get("inb"))
        self.s3 = (g.get("kn"), g.get("jn"), g.get("in"))

    def solve(self, parts=False, pp_residual=1e-6, pp_max_iter=2000):
        niter, resid, ms = C.c_int(), C.c_double(), C.c_float()
        assert self.lib.bbref_solve(1.0, 1e-3, pp_residual, pp_max_iter, int(parts), C.byref(niter), C.byref(resid), C.byref(ms)) == 0
        phi = np.zeros(self.s3b)
        assert self.lib.bbref_get(0, _P(phi)) == 0
        return niter.value, resid.value, phi[1:-1, 1:-1, 1:-1].copy()

    def get(self, which):
        a = np.zeros(self.s3 if which in (2, 3) else self.s3b)
        assert self.lib.bbref_get(which, _P(a)) == 0
        return a

    def spmv(self, vec, parts=False):
        out = np.zeros(self.s3)
        assert self.lib.bbref_spmv(_P(np.ascontiguousarray(vec)), int(parts), _P(out)) == 0
        return out

    def exchange(self, vec):
        a = np.ascontiguousarray(vec).copy()
        assert self.lib.bbref_exchange(_P(a)) == 0
        return a


@pytest.mark.parametrize("bc", ["cavity", "duct", "channel", "sedimentation", "periodic", "box"])
def test_solve_three_way(bc):
    from gpu_util import Product
    case = Case((32, 28, 36), bc=bc)
    ref = Ref(case)
    rn, rres, rphi = ref.solve()
    ores, ohist = case.solve_oracle()
    ophi = case.o.gather_interior(ob.PHI)
    p = Product(case)
    p.set_coefficients()
    res = p.solve()[0]
    phi = p.gather("phi")
    # oracle pinned by the reference
    assert ores.niter == rn
    assert rel_l2(ophi, rphi) < PHI_TOL
    assert abs(ores.resid - rres) <= 1e-6 * rres
    # product against the reference itself
    assert abs(res.niter - rn) <= 1
    if res.niter == rn:
        assert rel_l2(phi, rphi) < PHI_TOL
        assert abs(res.resid - rres) <= 1e-6 * rres
    # set-up pieces: rhs (PP_rhs) and the Jacobi diagonal
    assert np.abs(ref.get(1) - case.o.array(0, ob.RHS_P)).max() <= OP_TOL * np.abs(ref.get(1)).max()
    assert np.array_equal(ref.get(2), case.o.array(0, ob.INVM))
    p.close()


def test_solve_three_way_96_cavity():
    """BASELINE configs[0] (lid-driven cavity, 96^3, 1 rank) against the reference's kernels."""
    from gpu_util import Product
    case = Case((96, 96, 96), bc="cavity", omp=True)
    ref = Ref(case)
    rn, rres, rphi = ref.solve()
    ores, _ = case.solve_oracle()
    p = Product(case)
    p.set_coefficients()
    res = p.solve()[0]
    assert ores.niter == rn and res.niter == rn
    assert rel_l2(case.o.gather_interior(ob.PHI), rphi) < PHI_TOL
    assert rel_l2(p.gather("phi"), rphi) < PHI_TOL
    p.close()


def test_solve_three_way_particles():
    from gpu_util import Product
    case = Case((40, 40, 40), bc="sedimentation", nparts=4, radius=2.5)
    ref = Ref(case, nparts=4)
    rn, rres, rphi = ref.solve(parts=True)
    ores, _ = case.solve_oracle()
    p = Product(case)
    p.set_coefficients(parts=True)
    res = p.solve(parts=True)[0]
    assert ores.niter == rn and abs(res.niter - rn) <= 1
    assert rel_l2(case.o.gather_interior(ob.PHI), rphi) < PHI_TOL
    if res.niter == rn:
        assert rel_l2(p.gather("phi"), rphi) < PHI_TOL
    p.close()


@pytest.mark.parametrize("parts", [False, True])
def test_spmv_three_way(parts):
    from gpu_util import Product
    case = Case((32, 32, 32), bc="sedimentation", nparts=3 if parts else 0, radius=2.5)
    ref = Ref(case, nparts=3 if parts else 0)
    rng = np.random.default_rng(11)
    vec = rng.standard_normal(ref.s3b)
    rap = ref.spmv(vec, parts)
    case.o.array(0, ob.PB_Q)[...] = vec
    case.o.spmv(ob.PB_Q, parts=parts)
    oap = case.o.array(0, ob.APB_Q)
    p = Product(case)
    p.set_coefficients(parts=parts)
    gap = p.solvers[0].spmv(p.solvers[0].to_device(vec), use_phase=parts).cpu().numpy()
    scale = np.abs(rap).max()
    assert np.abs(oap - rap).max() <= OP_TOL * scale
    assert np.abs(gap - rap).max() <= OP_TOL * scale
    p.close()


@pytest.mark.parametrize("bc", ["periodic", "channel", "box"])
def test_exchange_three_way(bc):
    """cuda_BC_test_periodic analogue (src/cuda_testing.cu:749-933) with the reference's pack/unpack kernels."""
    from gpu_util import Product
    case = Case((16, 12, 20), bc=bc)
    ref = Ref(case)
    rng = np.random.default_rng(5)
    vec = rng.standard_normal(ref.s3b)
    rex = ref.exchange(vec)
    case.o.array(0, ob.PHI)[...] = vec
    case.o.exchange_Gcc(ob.PHI)
    assert np.array_equal(case.o.array(0, ob.PHI), rex)
    p = Product(case)
    p.dev[0]["phi"].copy_(p.solvers[0].to_device(vec))
    p.solvers[0].exchange_Gcc(p.dev[0]["phi"])
    assert np.array_equal(p.dev[0]["phi"].cpu().numpy(), rex)
    p.close()
