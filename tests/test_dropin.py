"""The drop-in layer driven the way Bluebottle drives it: tests/dropin_host.cu defines the reference's globals and
calls the reference's entry-point NAMES (void f(void)) in libbbpcg_dropin.so.  CPU: it links (every imported
global / callee resolves) and fails loudly without a GPU.  GPU: phi equals the oracle's, the solver_expd.rec line
is written through recorder_PP, non-convergence prints the reference's messages and exits with EXIT_FAILURE."""
import os
import subprocess

import numpy as np
import pytest

from cases import Case, rel_l2
from oracle import binding as ob
from test_domain import _write_flow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bluebottle-3.0_b200", "lib")
BUILD = os.path.join(ROOT, "tests", "_build")
EXE = os.path.join(BUILD, "dropin_host")



def _rec_lines(path):
    """(header tokens, tokens of the LAST line) of a record file written in the column format of src/recorder.c:157-221"""
    lines = open(path).read().split("\n")
    return lines[0].split(), lines[-1].split()

def _build():
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(ROOT, "tests", "dropin_host.cu")
    deps = [src, os.path.join(LIB, "libbbpcg.so"), os.path.join(LIB, "libbbpcg_dropin.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", EXE, src,
                           "-L" + LIB, "-lbbpcg_dropin", "-lbbpcg", "-Xlinker", "-rpath", "-Xlinker", LIB, "-lcudart"])
    return EXE


def _write_case(tmp, case, blocks=(1, 1, 1)):
    from bbpcg.grid import BC_SETS
    flow, dec, inp = str(tmp / "flow.config"), str(tmp / "decomp.config"), str(tmp / "inputs.bin")
    _write_flow(flow, case.extent, case.cells, blocks, BC_SETS[case.bcname])
    txt = open(flow).read().replace("rho_f 2.5", "rho_f 1.0").replace("pp_max_iter 1234", "pp_max_iter 2000").replace("pp_residual 1e-7", "pp_residual 1e-6")
    open(flow, "w").write(txt)
    import bbpcg
    bbpcg.Decomposition.uniform(case.extent, case.cells, blocks, BC_SETS[case.bcname]).write_decomp(dec, prec=6)
    a = case.inputs(0)
    with open(inp, "wb") as f:
        for k in ("flag_u", "flag_v", "flag_w", "phase", "phase_shell", "u_star", "v_star", "w_star"):
            f.write(np.ascontiguousarray(a[k]).tobytes())
    (tmp / "record").mkdir(exist_ok=True)
    return flow, dec, inp


def _write_case_ranks(tmp, case, blocks):
    """flow.config / decomp.config of the decomposition + one inputs.bin.<rank> per block"""
    flow, dec, inp = _write_case(tmp, case, blocks)
    os.remove(inp)
    for r in range(case.o.nblocks):
        a = case.inputs(r)
        with open("%s.%d" % (inp, r), "wb") as f:
            for k in ("flag_u", "flag_v", "flag_w", "phase", "phase_shell", "u_star", "v_star", "w_star"):
                f.write(np.ascontiguousarray(a[k]).tobytes())
    return flow, dec, inp


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="checks the behaviour WITHOUT a CUDA device")
def test_host_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build()
    case = Case((12, 10, 8), bc="duct")
    flow, dec, inp = _write_case(tmp_path, case)
    p = subprocess.run([exe, flow, dec, inp, str(tmp_path / "phi.bin"), str(tmp_path / "record"), "noparts"], capture_output=True, text=True)
    assert p.returncode == 1                                   # EXIT_FAILURE, the reference's error convention
    assert "bbpcg_create" in p.stderr and "no CUDA device" in p.stderr
    assert not os.path.exists(tmp_path / "phi.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("bc,parts", [("duct", False), ("cavity", False), ("sedimentation", True)])
def test_dropin_solve_matches_oracle(tmp_path, bc, parts):
    exe = _build()
    case = Case((32, 28, 36), bc=bc, nparts=3 if parts else 0, radius=2.5)
    flow, dec, inp = _write_case(tmp_path, case)
    out = str(tmp_path / "phi.bin")
    p = subprocess.run([exe, flow, dec, inp, out, str(tmp_path / "record"), "parts" if parts else "noparts"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    ores, _ = case.solve_oracle()
    g = case.o.dom(0).Gcc
    phi = np.fromfile(out, dtype=np.float64).reshape(g.get("knb"), g.get("jnb"), g.get("inb"))
    assert rel_l2(phi[1:-1, 1:-1, 1:-1], case.o.gather_interior(ob.PHI)) < 1e-10
    # mpi_cuda_exchange_Gcc(_phi) ran after the solve: ghosts equal the oracle's exchange of its own phi
    case.o.exchange_Gcc(ob.PHI)
    ophi = case.o.array(0, ob.PHI)
    for sl in ((slice(1, -1), slice(1, -1), 0), (slice(1, -1), slice(1, -1), -1), (slice(1, -1), 0, slice(1, -1)),
               (slice(1, -1), -1, slice(1, -1)), (0, slice(1, -1), slice(1, -1)), (-1, slice(1, -1), slice(1, -1))):
        assert np.allclose(phi[sl], ophi[sl], rtol=0, atol=1e-10 * np.abs(ophi).max())
    # recorder_PP received (niter, resid): one line in the reference's column format (src/recorder.c:190-221)
    head, rec = _rec_lines(tmp_path / "record" / "solver_expd.rec")
    assert head == ["stepnum", "ttime", "dt", "niter", "resid", "time", "(s)"]
    assert int(rec[0]) == 1 and int(rec[3]) == ores.niter and abs(float(rec[4]) - ores.resid) <= 1e-5 * ores.resid


@pytest.mark.gpu
def test_dropin_nonconvergence_prints_and_exits(tmp_path):
    """src/cuda_solver.cu:734-742: three printf lines and exit(EXIT_FAILURE) after pp_max_iter + 1 iterations"""
    exe = _build()
    case = Case((24, 24, 24), bc="periodic")
    flow, dec, inp = _write_case(tmp_path, case)
    p = subprocess.run([exe, flow, dec, inp, str(tmp_path / "phi.bin"), str(tmp_path / "record"), "noparts", "4"], capture_output=True, text=True)
    assert p.returncode == 1
    assert "The pressure-Poisson equation did not converge." in p.stdout
    assert "Residual at iteration 5 is" in p.stdout and "(rhs, rhs) is" in p.stdout
    assert not os.path.exists(tmp_path / "phi.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("bc", ["duct", "box"])
def test_dropin_epilogue_matches_oracle(tmp_path, bc):
    """cuda_dom_BC_p(_phi); cuda_project(); cuda_update_p() by their reference names (src/bluebottle.c:234-250)"""
    exe = _build()
    case = Case((20, 18, 22), bc=bc)
    flow, dec, inp = _write_case(tmp_path, case)
    out, epi = str(tmp_path / "phi.bin"), str(tmp_path / "epi.bin")
    p = subprocess.run([exe, flow, dec, inp, out, str(tmp_path / "record"), "noparts", "2000", epi], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    case.solve_oracle()
    d = case.o.dom(0)
    n = case.o.array(0, ob.P0).size
    C = np.arange(n, dtype=np.uint64)
    p0 = (((C * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(8)).astype(np.float64) / 16777216.0 - 0.5
    case.o.array(0, ob.P0)[...] = p0.reshape(case.o.array(0, ob.P0).shape)
    case.o.epilogue(1.0, 1e-3)
    raw = np.fromfile(epi, dtype=np.float64)
    off = 0
    for aid in (ob.U, ob.V, ob.W, ob.P):
        ref = case.o.array(0, aid)
        got = raw[off:off + ref.size].reshape(ref.shape)
        off += ref.size
        # phi comes from two different solvers (1e-10 relative L2): compare at the solve's accuracy
        assert np.abs(got - ref)[1:-1, 1:-1, 1:-1].max() <= 1e-8 * np.abs(ref).max(), aid
    assert off == raw.size


@pytest.mark.gpu
@pytest.mark.parametrize("blocks", [(1, 1, 2), (2, 1, 1)])
def test_dropin_two_processes_bootstrap_through_the_allgather_hook(tmp_path, blocks):
    """One process per rank, as under mpirun: each defines the reference's globals, the first entry point bootstraps the
    peer mapping through bb_dropin_allgather() (bbpcg_dropin.cu:46-54; MPI_Allgather in Bluebottle, files here), the solve
    and mpi_cuda_exchange_Gcc(_phi) then run over peer memory.  phi of both ranks against the oracle's decomposed solve."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs (two processes time-slicing ONE GPU would spin against each other)")
    exe = _build()
    case = Case((32, 28, 36), blocks=blocks, bc="duct")
    flow, dec, inp = _write_case_ranks(tmp_path, case, blocks)
    rdv = tmp_path / "rdv"
    rdv.mkdir()
    out = str(tmp_path / "phi.bin")
    procs = []
    for r in range(2):
        env = dict(os.environ, BB_RANK=str(r), BB_NPROCS="2", BB_RDV=str(rdv))
        procs.append(subprocess.Popen([exe, flow, dec, inp, out, str(tmp_path / "record"), "noparts"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    ores, _ = case.solve_oracle()
    case.o.exchange_Gcc(ob.PHI)
    for r in range(2):
        g = case.o.dom(r).Gcc
        phi = np.fromfile("%s.%d" % (out, r), dtype=np.float64).reshape(g.get("knb"), g.get("jnb"), g.get("inb"))
        ophi = case.o.array(r, ob.PHI)
        assert rel_l2(phi[1:-1, 1:-1, 1:-1], ophi[1:-1, 1:-1, 1:-1]) < 1e-10
        # ghosts from the NEIGHBOUR RANK's process (the exchange after the solve, src/bluebottle.c:233)
        for sl in ((slice(1, -1), slice(1, -1), 0), (slice(1, -1), slice(1, -1), -1), (0, slice(1, -1), slice(1, -1)), (-1, slice(1, -1), slice(1, -1))):
            assert np.allclose(phi[sl], ophi[sl], rtol=0, atol=1e-10 * np.abs(ophi).max())
        rec = _rec_lines(tmp_path / "record" / ("solver_expd.rec.%d" % r))[1]
        assert int(rec[3]) == ores.niter


@pytest.mark.gpu
def test_dropin_nan_prints_and_exits(tmp_path):
    """src/cuda_solver.cu:245-251: two printf lines naming the iteration and exit(EXIT_FAILURE)"""
    exe = _build()
    case = Case((24, 24, 24), bc="periodic")
    case.o.array(0, ob.U_STAR)[4, 5, 6] = np.nan
    flow, dec, inp = _write_case(tmp_path, case)
    p = subprocess.run([exe, flow, dec, inp, str(tmp_path / "phi.bin"), str(tmp_path / "record"), "noparts"], capture_output=True, text=True)
    assert p.returncode == 1
    assert "The PP equation did not converge." in p.stdout and "The residual at iteration 1 is nan" in p.stdout
    assert not os.path.exists(tmp_path / "phi.bin")


@pytest.mark.gpu
def test_dropin_converging_on_the_last_allowed_iteration_still_exits(tmp_path):
    """src/cuda_solver.cu:192,235-241,271-279: the loop runs q = 1 .. pp_max_iter + 1; converging ON iteration pp_max_iter + 1
    records the line, breaks, and still fails the `q > pp_max_iter` test"""
    exe = _build()
    case = Case((24, 24, 24), bc="periodic")
    ores, _ = case.solve_oracle()
    flow, dec, inp = _write_case(tmp_path, case)
    p = subprocess.run([exe, flow, dec, inp, str(tmp_path / "phi.bin"), str(tmp_path / "record"), "noparts", str(ores.niter - 1)],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "The pressure-Poisson equation did not converge." in p.stdout
    assert "Residual at iteration %d is" % ores.niter in p.stdout
    rec = _rec_lines(tmp_path / "record" / "solver_expd.rec")[1]
    assert int(rec[3]) == ores.niter                           # recorder_PP was still called (:239)
    p = subprocess.run([exe, flow, dec, inp, str(tmp_path / "phi.bin"), str(tmp_path / "record"), "noparts", str(ores.niter)],
                       capture_output=True, text=True)
    assert p.returncode == 0                                   # one more allowed iteration: a normal convergence


@pytest.mark.gpu
def test_dropin_timed_entry_point_writes_the_segment_log(tmp_path):
    """cuda_PP_cg_timed (src/cuda_solver.cu:302-571): same solve as cuda_PP_cg_noparts, logged to solver_expd_timed.rec with
    the eight segment columns of recorder_PP_timed (src/recorder.c:259-336); the fused kernels fill spmv and up1"""
    exe = _build()
    case = Case((32, 24, 40), bc="duct")
    ores, _ = case.solve_oracle()
    flow, dec, inp = _write_case(tmp_path, case)
    out = str(tmp_path / "phi.bin")
    p = subprocess.run([exe, flow, dec, inp, out, str(tmp_path / "record"), "timed"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    g = case.o.dom(0).Gcc
    phi = np.fromfile(out, dtype=np.float64).reshape(g.get("knb"), g.get("jnb"), g.get("inb"))
    assert rel_l2(phi[1:-1, 1:-1, 1:-1], case.o.array(0, ob.PHI)[1:-1, 1:-1, 1:-1]) < 1e-10
    lines = open(tmp_path / "record" / "solver_expd_timed.rec").read().split("\n")
    assert lines[0].split()[:5] == ["stepnum", "ttime", "dt", "niter", "resid"] and "spmv time (s)" in lines[0] and "mpi time (s)" in lines[0]
    rec = lines[-1].split()
    assert len(rec) == 14 and int(rec[3]) == ores.niter
    total, spmv, up1 = float(rec[5]), float(rec[6]), float(rec[9])
    assert spmv > 0. and up1 > 0. and spmv + up1 <= total          # device time of the two kernels inside the wall time of the call
    assert all(float(rec[i]) == 0. for i in (7, 8, 10, 11, 12, 13))
