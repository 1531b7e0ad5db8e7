#!/usr/bin/env python
"""Drive only the solve epilogue at bench size (for ncu): synthetic 512^3 duct block, 3 fused calls.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/epi_launches.csv python scripts/epi_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bbpcg  # noqa: E402
from bbpcg import synth  # noqa: E402
from bbpcg.grid import BC_SETS, grid_shape  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
dec = bbpcg.Decomposition.uniform((0., 12., 0., 12., 0., 12.), (g, g, g), (1, 1, 1), BC_SETS["duct"])
s = bbpcg.PoissonSolver(dec, 0, device=0)
dom = dec.doms[0]
fu, fv, fw = synth.flags_noparts_torch(dom, dec.DOM, dec.bc, dev)
u, v, w = synth.velocity_star_torch(dom, dec.DOM, dec.bc, dev)
phi = torch.rand(grid_shape(dom, "Gcc"), dtype=torch.float64, device=dev)
p0 = torch.rand(grid_shape(dom, "Gcc"), dtype=torch.float64, device=dev)
phase = torch.full(grid_shape(dom, "Gcc"), -1, dtype=torch.int32, device=dev)
un, vn, wn, pn = s.empty("Gfx"), s.empty("Gfy"), s.empty("Gfz"), s.empty("Gcc")
ms = [s.epilogue(phi, u, v, w, fu, fv, fw, un, vn, wn, p0, phase, pn) for _ in range(ncalls)]
print("epilogue ms:", ms, "GB/s at 104 B/cell:", [104 * g ** 3 / (m * 1e-3) / 1e9 for m in ms])
s.close()
