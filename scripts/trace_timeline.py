#!/usr/bin/env python
"""Per-CTA time line of the two iteration kernels from the debug build (make -C bluebottle-3.0_b200/csrc trace):
where a small-block iteration spends its time -- launch gap, ramp (first data), steady state, finishing spread, reduction tail.

    BBPCG_LIB_PATH=bluebottle-3.0_b200/lib/libbbpcg_trace.so python scripts/trace_timeline.py --grid 256 [--opt ty=6 --opt kc=86] --out gpurun_out/trace256.json
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "bluebottle-3.0_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="256")
    ap.add_argument("--bc", default="duct")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--opt", action="append", default=[])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "trace.json"))
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
    for kv in a.opt:
        k, x = kv.split("=")
        s.set_option(k, int(x))
    s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=20)
    r = s.PP_cg_noparts(u, v, w, rhs, phi, fixed_iters=a.iters)
    path = a.out + ".bin"
    s.lib.bbpcg_trace_dump.argtypes = [C.c_void_p, C.c_char_p]
    assert s.lib.bbpcg_trace_dump(s.h, path.encode()) == 0
    raw = np.fromfile(path, dtype=np.uint64)
    nl, nc, ne, launches = [int(x) for x in raw[:4]]
    t = raw[4:].reshape(nl, nc, ne).astype(np.int64)
    recs = []
    for slot in range(nl):
        used = t[slot, :, 3] > 0
        if not used.any():
            continue
        e = t[slot, used]
        kind = int(e[0, 7] & 0xff)
        launch = int(e[0, 7] >> 8)
        recs.append(dict(launch=launch, kind=kind, nctas=int(used.sum()), entry=e[:, 0], waited=e[:, 1], first=e[:, 2], loop_end=e[:, 3],
                         red=e[:, 4], end=int(e[:, 5].max()), sm=e[:, 6] & 0xffffffff, items=(e[:, 6] >> 32) & 0xff, planes=e[:, 6] >> 40))
    recs.sort(key=lambda x: x["launch"])
    recs = recs[2:-1]                                     # ring wrap: drop the partially overwritten oldest / the set-up neighbours
    out = {"cells": cells, "opts": a.opt, "us_per_iter": r.ms_iter * 1e3 / a.iters, "grid": s.info("search_grid"), "kc": s.info("search_kc"), "kernels": {}}
    prev_end = None
    rows = {1: [], 2: []}
    for x in recs:
        t0 = x["entry"].min()
        row = dict(gap_prev_end_to_first_entry=(t0 - prev_end) / 1e3 if prev_end else None,
                   entry_spread=(x["entry"].max() - t0) / 1e3,
                   pdl_wait_max=(x["waited"] - x["entry"]).max() / 1e3,
                   all_waited=(x["waited"].max() - t0) / 1e3,
                   first_data_median=float(np.median(x["first"] - x["waited"])) / 1e3,
                   loop_median=float(np.median(x["loop_end"] - x["first"])) / 1e3,
                   loop_min=float((x["loop_end"] - x["first"]).min()) / 1e3, loop_max=float((x["loop_end"] - x["first"]).max()) / 1e3,
                   first_cta_done=(x["loop_end"].min() - t0) / 1e3, last_cta_done=(x["loop_end"].max() - t0) / 1e3,
                   tail_after_last_cta=(x["end"] - x["loop_end"].max()) / 1e3,
                   # the tail split: every CTA's ticket (barrier + fence + atomic), the last CTA's item-order sum, then the scalars
                   ticket_median=float(np.median((x["red"] - x["loop_end"])[x["red"] > 0])) / 1e3,
                   last_cta_reduce=(x["red"].max() - x["loop_end"].max()) / 1e3,
                   scalars_after_reduce=(x["end"] - x["red"].max()) / 1e3,
                   total=(x["end"] - t0) / 1e3)
        rows[x["kind"]].append(row)
        prev_end = x["end"]
    for kind, name in ((1, "search"), (2, "resid")):
        if not rows[kind]:
            continue
        keys = rows[kind][0].keys()
        out["kernels"][name] = {k: float(np.median([r_[k] for r_ in rows[kind] if r_[k] is not None])) for k in keys}
    # CTAs per SM of the last search launch
    last = [x for x in recs if x["kind"] == 1][-1]
    sm, cnt = np.unique(last["sm"], return_counts=True)
    out["ctas_per_sm_hist"] = {int(k): int((cnt == k).sum()) for k in np.unique(cnt)}
    dur = (last["loop_end"] - last["first"]) / 1e3
    per_sm = {int(s_): int(c_) for s_, c_ in zip(sm, cnt)}
    solo = np.array([per_sm[int(s_)] == 1 for s_ in last["sm"]])
    out["loop_us_solo_sm"] = float(np.median(dur[solo])) if solo.any() else None
    out["loop_us_shared_sm"] = float(np.median(dur[~solo])) if (~solo).any() else None
    out["per_cta"] = {"sm": last["sm"].tolist(), "loop_us": [round(float(v), 2) for v in dur], "items": last["items"].tolist(), "planes": last["planes"].tolist(),
                      "end_us": [round(float(v), 2) for v in (last["loop_end"] - last["entry"].min()) / 1e3]}
    print(json.dumps({k: v for k, v in out.items() if k != "per_cta"}))
    with open(a.out, "a") as f:
        f.write(json.dumps(out) + "\n")
    os.remove(path)
    s.close()


if __name__ == "__main__":
    main()
