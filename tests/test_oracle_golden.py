"""CPU tests (no GPU): the oracle (oracle/pcg_ref.c, O2) against the committed golden vectors.

tests/golden/*.npz are OUTPUTS OF THE REFERENCE ITSELF: the unmodified reference kernels + host
loop (oracle/_ref/libbbref.so = /root/reference/src/{solver_kernel,cuda_solver,bluebottle_kernel}.cu)
run on a B200 by oracle/make_golden.py on the seeded inputs of tests/cases.py.  The reference ships
no golden vector for the Poisson solver, so these files are what pins the oracle.

Tolerances: integer / copy work (flags, phase, halo exchange, invM from 0/1 flags) is bit-exact;
one operator application agrees to FMA-contraction round-off (1e-13 of the field's max); the
solve agrees in iteration count exactly and in phi within 1e-10 relative L2 (north_star).
"""
import json
import os

import numpy as np
import pytest

from cases import Case, rel_l2
from oracle import binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "MANIFEST.json")) as f:
    MANIFEST = json.load(f)
CASES = MANIFEST["cases"]

PHI_TOL = 1e-10
OP_TOL = 1e-13


def _load(name):
    spec = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    case = Case(tuple(spec["cells"]), bc=spec["bc"], nparts=spec.get("nparts", 0), radius=spec.get("radius", 1.0))
    return spec, gold, case


def test_manifest_lists_every_fixture():
    files = sorted(f[:-4] for f in os.listdir(GOLD) if f.endswith(".npz") and not f.startswith(("epi_", "bcs_", "cage_")))   # epi_*: EPILOGUE_MANIFEST.json, bcs_*: BCSTAR_MANIFEST.json, cage_*: CAGES_MANIFEST.json
    assert files == sorted(CASES)
    assert MANIFEST["solve"] == {"rho_f": 1.0, "dt": 1e-3, "pp_residual": 1e-6, "pp_max_iter": 2000}


@pytest.mark.parametrize("name", sorted(CASES))
def test_inputs_are_the_ones_the_reference_saw(name):
    """the seeded inputs regenerate bit-identically (checksums recorded beside the outputs)"""
    spec, gold, case = _load(name)
    inp = case.inputs(0)
    assert float(sum(float(np.abs(inp[k]).sum()) for k in ("u_star", "v_star", "w_star"))) == float(gold["input_checksum"])
    assert int(sum(int(inp[k].sum()) for k in ("flag_u", "flag_v", "flag_w"))) == int(gold["flag_checksum"])
    if spec.get("nparts"):
        assert int((inp["phase"] > -1).sum()) == int(gold["phase_checksum"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_rhs_and_jacobi_diagonal(name):
    """PP_rhs (solver_kernel.cu:89-176) + the particle patch, and PP_jacobi_init (:26-87)"""
    spec, gold, case = _load(name)
    parts = bool(spec.get("nparts"))
    case.solve_oracle()                       # leaves rhs_p as the solve used it (patched when parts)
    rhs = case.o.array(0, ob.RHS_P)
    assert np.abs(rhs - gold["rhs"]).max() <= OP_TOL * np.abs(gold["rhs"]).max()
    assert np.array_equal(case.o.array(0, ob.INVM), gold["invM"])
    if parts:
        solid = case.o.array(0, ob.PHASE)[1:-1, 1:-1, 1:-1] > -1
        assert solid.any() and not rhs[1:-1, 1:-1, 1:-1][solid].any()       # particle_kernel.cu:1753


@pytest.mark.parametrize("name", sorted(CASES))
def test_spmv(name):
    """one application of -A on the seeded ghosted vector: PP_spmv_shared_load(_noparts)"""
    spec, gold, case = _load(name)
    g = case.o.dom(0).Gcc
    vec = np.random.default_rng(MANIFEST["spmv_vector_seed"]).standard_normal((g.get("knb"), g.get("jnb"), g.get("inb")))
    case.o.array(0, ob.PB_Q)[...] = vec
    case.o.spmv(ob.PB_Q, parts=False)
    ref = gold["ap_noparts"]
    assert np.abs(case.o.array(0, ob.APB_Q) - ref).max() <= OP_TOL * np.abs(ref).max()
    if spec.get("nparts"):
        case.o.spmv(ob.PB_Q, parts=True)
        ref = gold["ap_parts"]
        assert np.abs(case.o.array(0, ob.APB_Q) - ref).max() <= OP_TOL * np.abs(ref).max()
        # solid rows evaluate to minus identity
        solid = case.o.array(0, ob.PHASE)[1:-1, 1:-1, 1:-1] > -1
        assert np.allclose(case.o.array(0, ob.APB_Q)[solid], -vec[1:-1, 1:-1, 1:-1][solid], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_halo_exchange_bit_exact(name):
    """mpi_cuda_exchange_Gcc (mpi_comm.c:257-315) with the reference's pack/unpack kernels"""
    spec, gold, case = _load(name)
    g = case.o.dom(0).Gcc
    vec = np.random.default_rng(MANIFEST["spmv_vector_seed"]).standard_normal((g.get("knb"), g.get("jnb"), g.get("inb")))
    case.o.array(0, ob.PHI)[...] = vec
    case.o.exchange_Gcc(ob.PHI)
    assert np.array_equal(case.o.array(0, ob.PHI), gold["exchanged"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_solve(name):
    """cuda_PP_cg / cuda_PP_cg_noparts (cuda_solver.cu:38-300, 573-761): iterations, residual, phi"""
    spec, gold, case = _load(name)
    res, hist = case.solve_oracle()
    assert res.status == 0
    assert res.niter == int(gold["niter"])
    assert abs(res.resid - float(gold["resid"])) <= 1e-6 * float(gold["resid"])
    assert rel_l2(case.o.gather_interior(ob.PHI), gold["phi"]) < PHI_TOL
    assert len(hist) == res.niter + 1 and hist[-1] <= 1e-12 * res.sp_rhs and np.all(hist[:-1] > 1e-12 * res.sp_rhs)


@pytest.mark.parametrize("blocks", [(2, 1, 1), (1, 2, 1), (1, 1, 2), (2, 2, 1), (2, 2, 2), (3, 1, 2)])
def test_decomposed_oracle_reproduces_the_single_block_golden(blocks):
    """The discrete solution is decomposition independent: the multi-block oracle
    (halo exchange + rank-ordered all-reduce restated on the CPU) must land on the reference's
    1-rank answer."""
    name = "cavity_24x20x28" if blocks != (3, 1, 2) else "periodic_24x20x28"
    spec = CASES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    case = Case(tuple(spec["cells"]), blocks=blocks, bc=spec["bc"])
    res, _ = case.solve_oracle()
    assert abs(res.niter - int(gold["niter"])) <= 1
    if res.niter == int(gold["niter"]):
        assert rel_l2(case.o.gather_interior(ob.PHI), gold["phi"]) < PHI_TOL


def test_omp_oracle_is_bit_identical_to_serial():
    a = Case((24, 20, 28), bc="duct")
    b = Case((24, 20, 28), bc="duct", omp=True)
    ra, ha = a.solve_oracle()
    rb, hb = b.solve_oracle()
    assert ra.niter == rb.niter and np.array_equal(ha, hb)
    assert np.array_equal(a.o.array(0, ob.PHI), b.o.array(0, ob.PHI))
