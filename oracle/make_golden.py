#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE'S OWN kernels (oracle/_ref/libbbref.so:
/root/reference/src/{solver_kernel,cuda_solver,bluebottle_kernel}.cu compiled unmodified, see
oracle/Makefile and oracle/ref_shim.cu) on seeded synthetic inputs.  Needs a GPU:

    gpurun -- python oracle/make_golden.py gpurun_out/golden      # then copy into tests/golden/

The reference ships no golden vector, known-answer test or fixture for the Poisson solver
(SURVEY.md 4, 8c), so these files -- outputs of the reference itself -- are what pins the CPU
oracle (tests/test_oracle_golden.py, no GPU needed) and, through it, the CUDA product.
TEST INFRASTRUCTURE ONLY.  Each case runs in its own process (the reference keeps its state in
globals: one grid per process).
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = {
    # name: dict(cells, bc, nparts, radius)
    "cavity_24x20x28": dict(cells=(24, 20, 28), bc="cavity"),
    "duct_24x20x28": dict(cells=(24, 20, 28), bc="duct"),
    "channel_24x20x28": dict(cells=(24, 20, 28), bc="channel"),
    "sedimentation_24x20x28": dict(cells=(24, 20, 28), bc="sedimentation"),
    "periodic_24x20x28": dict(cells=(24, 20, 28), bc="periodic"),
    "box_24x20x28": dict(cells=(24, 20, 28), bc="box"),
    "cavity_40": dict(cells=(40, 40, 40), bc="cavity"),                      # > 50 iterations: crosses the q%50 refresh
    "ragged_33x17x9": dict(cells=(33, 17, 9), bc="channel"),
    "parts_32": dict(cells=(32, 32, 32), bc="sedimentation", nparts=3, radius=2.5),
    "parts_40_duct": dict(cells=(40, 40, 40), bc="duct", nparts=4, radius=2.5),
}


def run_case(name, outdir):
    import numpy as np
    from cases import Case, load_ref
    from oracle import binding as ob
    spec = CASES[name]
    nparts = spec.get("nparts", 0)
    case = Case(spec["cells"], bc=spec["bc"], nparts=nparts, radius=spec.get("radius", 1.0))
    lib = load_ref()
    assert lib is not None, "oracle/_ref/libbbref.so missing"
    dom, DOM = case.o.dom(0), case.o.DOM
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    inp = case.inputs(0)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    keep = {k: np.ascontiguousarray(v) for k, v in inp.items()}
    assert lib.bbref_set_inputs(P(keep["flag_u"]), P(keep["flag_v"]), P(keep["flag_w"]), P(keep["phase"]),
                                P(keep["phase_shell"]), P(keep["u_star"]), P(keep["v_star"]), P(keep["w_star"]), nparts) == 0
    g = dom.Gcc
    s3b_shape = (g.get("knb"), g.get("jnb"), g.get("inb"))
    s3_shape = (g.get("kn"), g.get("jn"), g.get("in"))
    rng = np.random.default_rng(11)
    vec = rng.standard_normal(s3b_shape)
    # (1) one operator application on a seeded ghosted vector (both operators when particles exist)
    ap_noparts = np.zeros(s3_shape)
    # invM first: the SpMV unit call needs nothing from it, but bbref_solve computes it
    niter, resid, ms = C.c_int(), C.c_double(), C.c_float()
    assert lib.bbref_solve(1.0, 1e-3, 1e-6, 2000, 1 if nparts else 0, C.byref(niter), C.byref(resid), C.byref(ms)) == 0
    phi, rhs, invM = np.zeros(s3b_shape), np.zeros(s3b_shape), np.zeros(s3_shape)
    assert lib.bbref_get(0, P(phi)) == 0 and lib.bbref_get(1, P(rhs)) == 0 and lib.bbref_get(2, P(invM)) == 0
    assert lib.bbref_spmv(P(vec), 0, P(ap_noparts)) == 0
    out = dict(niter=np.int64(niter.value), resid=np.float64(resid.value), phi=phi[1:-1, 1:-1, 1:-1].copy(), rhs=rhs, invM=invM,
               ap_noparts=ap_noparts, input_checksum=np.float64(sum(float(np.abs(keep[k]).sum()) for k in ("u_star", "v_star", "w_star"))),
               flag_checksum=np.int64(sum(int(keep[k].sum()) for k in ("flag_u", "flag_v", "flag_w"))))
    if nparts:
        ap_parts = np.zeros(s3_shape)
        assert lib.bbref_spmv(P(vec), 1, P(ap_parts)) == 0
        out["ap_parts"] = ap_parts
        out["phase_checksum"] = np.int64(int((keep["phase"] > -1).sum()))
    # (2) the Gcc halo exchange (pack / self-put / unpack kernels) on the same seeded vector
    ex = vec.copy()
    assert lib.bbref_exchange(P(ex)) == 0
    # store only the ghost shell difference compactly: the full array is small enough
    out["exchanged"] = ex
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
    print(name, "niter", niter.value, "resid %.6e" % resid.value, "ms %.2f" % ms.value)


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2:
        run_case(sys.argv[2], outdir)
        return
    for name in CASES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), outdir, name])
    with open(os.path.join(outdir, "MANIFEST.json"), "w") as f:
        json.dump({"cases": CASES, "generator": "oracle/make_golden.py", "source": "reference kernels via oracle/_ref/libbbref.so on B200",
                   "solve": {"rho_f": 1.0, "dt": 1e-3, "pp_residual": 1e-6, "pp_max_iter": 2000}, "spmv_vector_seed": 11}, f, indent=1)


if __name__ == "__main__":
    main()
