#!/usr/bin/env python
"""Tuning sweep on ONE GPU: iteration-loop time of a fixed-iteration solve for a list of
(tile, kc, ...) settings of bbpcg_set_option.  Writes one JSON line per setting.

    python scripts/sweep.py --grid 256 --opt kc=16,32,58,64,128,256 --opt tile=0,1 [--iters 100] [--out gpurun_out/sweep.jsonl]
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="512")
    ap.add_argument("--bc", default="duct")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--repeat", type=int, default=1, help="passes over the whole list of settings (interleaved: box drift hits every setting alike); each line carries its pass")
    a = ap.parse_args()
    import torch
    import bbpcg
    from bbpcg import synth
    from bbpcg.grid import BC_SETS
    g = [int(v) for v in a.grid.split(",")]
    cells = tuple(g) if len(g) == 3 else (g[0],) * 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dec = bbpcg.Decomposition.uniform((0., 12., 0., 12. * cells[1] / cells[0], 0., 12. * cells[2] / cells[0]), cells, (1, 1, 1), BC_SETS[a.bc])
    s = bbpcg.PoissonSolver(dec, 0, device=0)
    dom = dec.doms[0]
    s.init_jacobi_preconditioner(*synth.flags_noparts_torch(dom, dec.DOM, dec.bc, dev))
    u, v, w = synth.velocity_star_torch(dom, dec.DOM, dec.bc, dev)
    rhs, phi = s.empty("Gcc"), s.empty("Gcc")
    keys = [kv.split("=")[0] for kv in a.opt]
    vals = [[int(x) for x in kv.split("=")[1].split(",")] for kv in a.opt]
    ncell = cells[0] * cells[1] * cells[2]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=20)
    with open(a.out, "a") as f:
        combos = list(itertools.product(*vals)) if vals else [()]
        for rep, combo in [(rep, c) for rep in range(a.repeat) for c in combos]:
            for k, x in zip(keys, combo):
                s.set_option(k, x)
            s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=10)
            s.set_option("kernel_timing", 1)
            r = s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=a.iters)
            s.set_option("kernel_timing", 0)
            r2 = s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=a.iters)
            us = r2.ms_iter * 1e3 / a.iters
            rec = {"cells": cells, "opts": dict(zip(keys, combo)), "pass": rep, "us_per_iter": us, "its": 1e6 / us,
                   "frac72": 72 * ncell / (us * 1e-6) / 1e9 / 6543.1,
                   "search_us": s.info("kt_search_ns") / max(s.info("kt_search_n"), 1) / 1e3,
                   "resid_us": s.info("kt_resid_ns") / max(s.info("kt_resid_n"), 1) / 1e3,
                   "refresh_us": s.info("kt_refresh_ns") / max(s.info("kt_refresh_n"), 1) / 1e3,
                   "search_grid": s.info("search_grid"), "search_kc": s.info("search_kc")}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")
    s.close()


if __name__ == "__main__":
    main()
