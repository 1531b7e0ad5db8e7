#!/usr/bin/env python
"""The workload compute-sanitizer runs (scripts/gpu_sanitize.sh): every kernel of the library once or more on small grids --
a particle-free solve crossing the q % 50 refresh on a ragged periodic-x grid (general plane loops, self-wrap halo pulls
through the x-face buffers and the y/z ghost tiles), an XFULL grid (in = 128), a particle solve, the four halo exchanges,
cuda_solvability and the epilogue.  `--ranks 2`: the same on a 2 x 1 x 1 decomposition driven from this one process.
Prints SANITIZE_CASE_OK when the results agree with the CPU oracle."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=1)
ap.add_argument("--small", action="store_true", help="racecheck is ~100x slower: one tiny solve only")
a = ap.parse_args()

import numpy as np  # noqa: E402
from cases import Case, rel_l2  # noqa: E402
from gpu_util import Product  # noqa: E402
from oracle import binding as ob  # noqa: E402

blocks = (2, 1, 1) if a.ranks == 2 else (1, 1, 1)


def solve(case, parts=False, **opts):
    p = Product(case, options=dict(opts, comm_timeout_ms=120000))
    p.set_coefficients(parts=parts)
    res = p.solve(parts=parts)
    ores, _ = case.solve_oracle()
    err = rel_l2(p.gather("phi"), case.o.gather_interior(ob.PHI))
    print("solve %s blocks %s parts %d: niter %d (oracle %d) rel-L2 %.2e" % (case.cells, blocks, parts, res[0].niter, ores.niter, err), flush=True)
    assert res[0].niter == ores.niter and err < 1e-10
    return p


if a.small:
    solve(Case((20, 12, 10), blocks=blocks, bc="duct"), kc=4).close()
else:
    p = solve(Case((40, 24, 36), blocks=blocks, bc="duct"), kc=9)              # > 50 iterations: refresh kernels too
    s, d = p.solvers[0], p.dev[0]
    for grid, key in (("Gcc", "phi"), ("Gfx", "u_star"), ("Gfy", "v_star"), ("Gfz", "w_star")):
        p.each(lambda r, s_, d_: s_.exchange(d_[key], grid))
    p.each(lambda r, s_, d_: s_.solvability(d_["u_star"], d_["v_star"], d_["w_star"], "HOMOGENEOUS"))
    case = p.case
    case.seed_epilogue(5, phi=False)

    def epi(r, s_, d_):
        d_["p0"] = s_.to_device(case.o.array(r, ob.P0))
        out = [s_.empty(g) for g in ("Gfx", "Gfy", "Gfz", "Gcc")]
        s_.epilogue(d_["phi"], d_["u_star"], d_["v_star"], d_["w_star"], d_["flag_u"], d_["flag_v"], d_["flag_w"], out[0], out[1], out[2],
                    d_["p0"], d_["phase"], out[3])
        return out
    p.each(epi)
    p.close()
    solve(Case((128, 12, 14), blocks=(1, 1, 1) if a.ranks == 1 else (1, 2, 1), bc="channel"), ty=7, kc=5).close()    # XFULL plane loops
    ext = (0., 12., 0., 12., 0., 12.)
    solve(Case((36, 36, 36), blocks=blocks, bc="sedimentation", extent=ext, nparts=3, radius=2.5), parts=True).close()
print("SANITIZE_CASE_OK", flush=True)
