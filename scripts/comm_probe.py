#!/usr/bin/env python
"""Where does the exposed communication time of a decomposed solve go?  (torchrun, N >= 2)
Per-kernel CUDA-event times (kernel_timing = 1: no PDL) of the decomposed solve vs the same per-rank block solved
stand-alone, plus loop times with PDL on/off.  Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bbpcg  # noqa: E402
from bbpcg import synth  # noqa: E402
from bbpcg.grid import BC_SETS  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
blocks = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,1,2").split(","))
cells = tuple(int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "512,512,512").split(","))
g = cells[0]
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
NIT = 200


def mx(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def setup(dec, r, attach):
    s = bbpcg.PoissonSolver(dec, r, device=local)
    if attach:
        s.comm_init_torch()
    d = dec.doms[r]
    s.init_jacobi_preconditioner(*synth.flags_noparts_torch(d, dec.DOM, dec.bc, dev))
    u, v, w = synth.velocity_star_torch(d, dec.DOM, dec.bc, dev)
    return s, (u, v, w, s.empty("Gcc"), s.empty("Gcc"))


def run(s, arrs, **opts):
    for k, v in opts.items():
        s.set_option(k, v)
    s.PP_cg_noparts(*arrs, fixed_iters=20)
    dist.barrier(); torch.cuda.synchronize()
    r = s.PP_cg_noparts(*arrs, fixed_iters=NIT)
    out = {"loop_us": mx(r.ms_iter * 1e3 / NIT)}
    if opts.get("kernel_timing"):
        out["search_us"] = mx(s.info("kt_search_ns") * 1e-3 / max(s.info("kt_search_n"), 1))
        out["resid_us"] = mx(s.info("kt_resid_ns") * 1e-3 / max(s.info("kt_resid_n"), 1))
        out["refresh_us"] = mx(s.info("kt_refresh_ns") * 1e-3 / max(s.info("kt_refresh_n"), 1))
    return out


L = 12.0
dec = bbpcg.Decomposition.uniform((0., L, 0., L * cells[1] / g, 0., L * cells[2] / g), cells, blocks, BC_SETS["duct"])
dom = dec.doms[rank]
dec1 = bbpcg.Decomposition.uniform((dom.xs, dom.xe, dom.ys, dom.ye, dom.zs, dom.ze), (dom.xn, dom.yn, dom.zn), (1, 1, 1), BC_SETS["duct"])
sN, aN = setup(dec, rank, True)
s1, a1 = setup(dec1, 0, False)
res = {"blocks": blocks, "cells": cells, "world": world}
res["multi_pdl1"] = run(sN, aN, pdl=1, kernel_timing=0)
res["alone_pdl1"] = run(s1, a1, pdl=1, kernel_timing=0)
res["multi_pdl0"] = run(sN, aN, pdl=0, kernel_timing=0)
res["alone_pdl0"] = run(s1, a1, pdl=0, kernel_timing=0)
res["multi_timed"] = run(sN, aN, pdl=1, kernel_timing=1)
res["alone_timed"] = run(s1, a1, pdl=1, kernel_timing=1)
if rank == 0:
    print("COMM_PROBE " + json.dumps(res))
sN.close(); s1.close()
dist.barrier()
dist.destroy_process_group()
