#!/usr/bin/env python
"""Generate tests/golden/epi_*.npz: outputs of the REFERENCE'S OWN epilogue kernels (BC_p_*_N, project_u/v/w,
update_p, copy_p_p_noghost, forcing_add_c_const from /root/reference/src/bluebottle_kernel.cu, compiled unmodified
into oracle/_ref/libbbref.so and driven in the order of src/bluebottle.c:233-250) on seeded inputs.  Needs a GPU:

    gpurun -- python oracle/make_golden_epilogue.py gpurun_out/golden_epi     # then copy into tests/golden/

The reference has no test for these kernels; these files pin the CPU restatement
(oracle/pcg_ref.c: bbo_dom_BC_p / bbo_project / bbo_update_p) in tests/test_oracle_epilogue.py without a GPU.
TEST INFRASTRUCTURE ONLY.  One process per case (the reference keeps its state in globals).
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
    "epi_cavity_12x10x14": dict(cells=(12, 10, 14), bc="cavity"),
    "epi_duct_12x10x14": dict(cells=(12, 10, 14), bc="duct"),
    "epi_periodic_12x10x14": dict(cells=(12, 10, 14), bc="periodic"),
    "epi_box_ragged_19x7x17": dict(cells=(19, 7, 17), bc="box"),          # not multiples of the 16-cell tile
    "epi_parts_16": dict(cells=(16, 16, 16), bc="sedimentation", nparts=1, radius=3.0),
}
SEED = 23


def run_case(name, outdir):
    import numpy as np
    from cases import Case, load_ref, ref_epilogue
    spec = CASES[name]
    nparts = spec.get("nparts", 0)
    case = Case(spec["cells"], bc=spec["bc"], nparts=nparts, radius=spec.get("radius", 1.0))
    case.seed_epilogue(SEED)
    lib = load_ref()
    assert lib is not None, "oracle/_ref/libbbref.so missing"
    dom, DOM = case.o.dom(0), case.o.DOM
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    keep = {k: np.ascontiguousarray(v) for k, v in case.inputs(0).items()}
    assert lib.bbref_set_inputs(P(keep["flag_u"]), P(keep["flag_v"]), P(keep["flag_w"]), P(keep["phase"]),
                                P(keep["phase_shell"]), P(keep["u_star"]), P(keep["v_star"]), P(keep["w_star"]), nparts) == 0
    ein = case.epilogue_inputs(0)
    out = ref_epilogue(lib, case, ein["phi"], ein["p0"])
    # the face-grid halo exchanges (pack / self-put / unpack kernels of Gfx, Gfy, Gfz) on seeded arrays
    from cases import face_exchange_inputs
    ex = {}
    for key, (arr, code) in face_exchange_inputs(case, 0, SEED + 6).items():
        a = np.ascontiguousarray(arr).copy()
        assert lib.bbref_exchange_face(P(a), code) == 0
        ex["ex_" + key] = a
    # cuda_solvability with the reference's surf_int_* / plane_eps_* kernels on the same seeded arrays
    for out_plane in (10, 1, 4):
        arrs = {k: np.ascontiguousarray(v[0]).copy() for k, v in face_exchange_inputs(case, 0, SEED + 6).items()}
        eps = (C.c_double * 3)()
        assert lib.bbref_solvability(P(arrs["u"]), P(arrs["v"]), P(arrs["w"]), out_plane, eps) == 0
        ex["sol_eps_%d" % out_plane] = np.array([eps[0], eps[1], eps[2]])
        fresh = {k: v[0] for k, v in face_exchange_inputs(case, 0, SEED + 6).items()}
        for k in "uvw":                                       # only the arrays the reference changed are stored
            if out_plane == 10 or not np.array_equal(arrs[k], fresh[k]):
                ex["sol_%s_%d" % (k, out_plane)] = arrs[k]
            else:
                ex["sol_same_%s_%d" % (k, out_plane)] = np.int8(1)
    np.savez_compressed(os.path.join(outdir, name + ".npz"), u=out["u"], v=out["v"], w=out["w"], p=out["p"], phi=out["phi"], **ex,
                        input_checksum=np.float64(float(np.abs(ein["phi"]).sum() + np.abs(ein["p0"]).sum())))
    print(name, "ms %.3f" % out["ms"], "mean(p) %.3e" % out["p"][1:-1, 1:-1, 1:-1].mean())


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_epi")
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2:
        run_case(sys.argv[2], outdir)
        return
    for name in CASES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), outdir, name])
    with open(os.path.join(outdir, "EPILOGUE_MANIFEST.json"), "w") as f:
        json.dump({"cases": CASES, "generator": "oracle/make_golden_epilogue.py", "seed": SEED,
                   "source": "reference kernels via oracle/_ref/libbbref.so on B200", "rho_f": 1.0, "dt": 1e-3}, f, indent=1)


if __name__ == "__main__":
    main()
