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

pos = [x for x in sys.argv[1:] if "=" not in x]
opts = dict((x.split("=")[0], int(x.split("=")[1])) for x in sys.argv[1:] if "=" in x)      # bbpcg_set_option pairs: epi_chunk=64 ...
g = int(pos[0]) if len(pos) > 0 else 512
ncalls = int(pos[1]) if len(pos) > 1 else 3
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
for k_, v_ in opts.items():
    s.set_option(k_, v_)
ms = [s.epilogue(phi, u, v, w, fu, fv, fw, un, vn, wn, p0, phase, pn) for _ in range(ncalls)]
print("epilogue", opts, "ms:", ["%.3f" % m for m in ms], "GB/s at 104 B/cell:", [104 * g ** 3 / (m * 1e-3) / 1e9 for m in ms])
s.close()
