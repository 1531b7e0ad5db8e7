#!/usr/bin/env python
"""Generate tests/golden/cage_*.npz: outputs of the REFERENCE'S OWN cage kernels -- reset_flag_*, reset_phases, cage_setup,
build_phase, build_phase_shell, cage_flag_*, flag_external_* (/root/reference/src/particle_kernel.cu:79-576, compiled
unmodified into oracle/_ref/libbbref.so) driven in the order of cuda_build_cages (src/cuda_particle.cu:1516-1646).  Needs a GPU:

    gpurun -- python oracle/make_golden_cages.py gpurun_out/golden_cages      # then copy into tests/golden/

The reference has no test for these kernels; the files pin oracle/pcg_ref.c: bbo_build_cages in tests/test_cages.py without
a GPU.  TEST INFRASTRUCTURE ONLY.  One process per case (the reference keeps its state in globals).
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
    "cage_inside_sedimentation_24x20x28": dict(cells=(24, 20, 28), bc="sedimentation", parts="inside"),
    "cage_faces_sedimentation_24x20x28": dict(cells=(24, 20, 28), bc="sedimentation", parts="faces"),   # x, y periodic (cage reaches the ghosts), z walls (clipped to the interior)
    "cage_faces_box_21x17x19": dict(cells=(21, 17, 19), bc="box", parts="faces"),                       # odd cell counts: cage_dim loses a cell (particle_kernel.cu:142-145)
    "cage_inside_periodic_16": dict(cells=(16, 16, 16), bc="periodic", parts="inside"),
    "cage_none_duct_12x10x14": dict(cells=(12, 10, 14), bc="duct", parts=None),                         # NPARTS == 0: flags only
}


def run_case(name, outdir):
    import numpy as np
    from cases import Case, cage_particles, load_ref, ref_build_cages
    spec = CASES[name]
    case = Case(spec["cells"], bc=spec["bc"])
    lib = load_ref()
    assert lib is not None, "oracle/_ref/libbbref.so missing"
    dom, DOM = case.o.dom(0), case.o.DOM
    assert lib.bbref_init(C.byref(dom), C.byref(DOM)) == 0
    parts = cage_particles(spec["parts"], case.extent, case.cells) if spec["parts"] else tuple(np.zeros(0) for _ in range(4))
    out = ref_build_cages(lib, case, parts)
    if not spec["parts"]:
        out.pop("phase"); out.pop("phase_shell")               # untouched by the reference when NPARTS == 0
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
    print(name, "ok", {k: int((v != 1).sum()) for k, v in out.items()})


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_cages")
    os.makedirs(outdir, exist_ok=True)
    if len(sys.argv) > 2:
        run_case(sys.argv[2], outdir)
        return
    for name in CASES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), outdir, name])
    with open(os.path.join(outdir, "CAGES_MANIFEST.json"), "w") as f:
        json.dump({"cases": CASES, "generator": "oracle/make_golden_cages.py",
                   "source": "reference kernels via oracle/_ref/libbbref.so on B200"}, f, indent=1)


if __name__ == "__main__":
    main()
