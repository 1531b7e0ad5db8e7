#!/usr/bin/env python
"""Strong-scaling probe (torchrun, N ranks): the same global grid under several block decompositions, PDL on/off,
fixed iteration count; rank 0 prints one JSON line per configuration (loop us/iteration = max over ranks)."""
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
decomps = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1].split(";")]
g = int(sys.argv[2]) if len(sys.argv) > 2 else 512
bc = sys.argv[3] if len(sys.argv) > 3 else "duct"
NIT = 200
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def mx(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for blocks in decomps:
    dec = bbpcg.Decomposition.uniform((0., 12., 0., 12., 0., 12.), (g, g, g), blocks, BC_SETS[bc])
    s = bbpcg.PoissonSolver(dec, rank, device=local)
    s.comm_init_torch()
    d = dec.doms[rank]
    s.init_jacobi_preconditioner(*synth.flags_noparts_torch(d, dec.DOM, dec.bc, dev))
    arrs = list(synth.velocity_star_torch(d, dec.DOM, dec.bc, dev)) + [s.empty("Gcc"), s.empty("Gcc")]
    out = {"blocks": blocks, "grid": g, "world": world, "bc": bc, "block_cells": (d.xn, d.yn, d.zn)}
    for pdl in (1, 0, 1, 0):
        s.set_option("pdl", pdl)
        s.PP_cg_noparts(*arrs, fixed_iters=20)
        dist.barrier(); torch.cuda.synchronize()
        r = s.PP_cg_noparts(*arrs, fixed_iters=NIT)
        out.setdefault("loop_us_pdl%d" % pdl, []).append(round(mx(r.ms_iter * 1e3 / NIT), 1))
    r = s.PP_cg_noparts(*arrs)                      # one converged solve with the default options
    out["converged_iters"], out["solve_ms"] = r.niter, round(mx(r.ms_total), 2)
    if rank == 0:
        print("DECOMP_PROBE " + json.dumps(out), flush=True)
    s.close()
    del arrs
    dist.barrier()
dist.destroy_process_group()
