#!/usr/bin/env python
"""Sustained vs burst HBM copy bandwidth on this box (what MEASURED_PEAKS.json's hbm_gbs is the burst figure of):
b.copy_(a) over 1 Gi bf16 elements, best of 10 (burst) and back to back for ~4 s (sustained), with the SM clock
sampled during the sustained leg.  One JSON line."""
import json
import subprocess
import time

import torch

torch.cuda.set_device(0)
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
a.fill_(1.0)
nbytes = 2 * n * 2
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                     stdout=subprocess.PIPE, text=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
reps = 0
while time.perf_counter() - t0 < 4.0:
    for _ in range(20):
        b.copy_(a)
    reps += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
p.terminate()
lines = [ln.split(",") for ln in p.stdout.read().strip().splitlines()]
sm = sorted(float(x[0]) for x in lines if len(x) == 2)
print(json.dumps({"hbm_copy_burst_gbs": best, "hbm_copy_sustained_gbs": sus, "reps": reps,
                  "sm_mhz_median_sustained": sm[len(sm) // 2] if sm else None}))
