#!/usr/bin/env python
"""Generate tests/golden/bcs_*.npz: outputs of the REFERENCE'S OWN velocity boundary-condition kernels
BC_{u,v,w}_{W,E,S,N,B,T}_{D,N} (/root/reference/src/bluebottle_kernel.cu:104-598, compiled unmodified into
oracle/_ref/libbbref.so) driven by the switch table of cuda_dom_BC_star (src/cuda_bluebottle.cu:2111-2311) on seeded
u*, v*, w*.  Needs a GPU:

    gpurun -- python oracle/make_golden_bcstar.py gpurun_out/golden_bcs      # then copy into tests/golden/

The reference has no test for these kernels; the files pin oracle/pcg_ref.c: bbo_dom_BC_star in tests/test_bc_star.py
without a GPU.  TEST INFRASTRUCTURE ONLY.  One process per case (the reference keeps its state in globals).
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

# pressure BC set decides which sides have no neighbour (periodic sides wrap onto the block itself: no BC there)
CASES = {
    "bcs_box_12x10x14": dict(cells=(12, 10, 14), bc="box"),
    "bcs_duct_9x7x5": dict(cells=(9, 7, 5), bc="duct"),
    "bcs_cavity_thin_2x1x3": dict(cells=(2, 1, 3), bc="cavity"),       # one- and two-cell-thick directions: W then E on the same line
}
TABLES = ("dirichlet", "neumann", "cavity_lid", "mixed")
SEED = 41


def run_case(name, outdir):
    import numpy as np
    from cases import Case, face_exchange_inputs, load_ref, ref_dom_BC_star
    spec = CASES[name]
    case = Case(spec["cells"], bc=spec["bc"])
    lib = load_ref()
    assert lib is not None, "oracle/_ref/libbbref.so missing"
    dom, DOM = case.o.dom(0), case.o.DOM
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    out = {}
    for t in TABLES:
        arrs = {k: np.ascontiguousarray(v[0]).copy() for k, v in face_exchange_inputs(case, 0, SEED).items()}
        ref_dom_BC_star(lib, case, arrs, t)
        for k in "uvw":
            out["%s_%s" % (t, k)] = arrs[k]
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
    print(name, "ok")


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_bcs")
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2:
        run_case(sys.argv[2], outdir)
        return
    for name in CASES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), outdir, name])
    from cases import BC_STAR_TABLES
    with open(os.path.join(outdir, "BCSTAR_MANIFEST.json"), "w") as f:
        json.dump({"cases": CASES, "tables": {t: BC_STAR_TABLES[t] for t in TABLES}, "generator": "oracle/make_golden_bcstar.py",
                   "seed": SEED, "source": "reference kernels via oracle/_ref/libbbref.so on B200"}, f, indent=1)


if __name__ == "__main__":
    main()
