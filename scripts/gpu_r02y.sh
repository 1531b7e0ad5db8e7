#!/bin/bash
# round 2, call Y (1 GPU): guided-plan sweep, 4 interleaved passes (call X drifted by 5 % from its first to its last setting)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02y_sweep.jsonl
for g in 256,256,256 512,256,128; do
  timeout 400 python scripts/sweep.py --grid $g --iters 150 --repeat 4 --opt pdl=1 --opt chunk_min=4,8 --opt guided_pct=60,100 --opt ty=7,8 --out gpurun_out/r02y_sweep.jsonl > /dev/null 2>> gpurun_out/r02y_sweep.err
done
tail -2 gpurun_out/r02y_sweep.err
python - <<'PY'
import json, collections
acc = collections.defaultdict(list)
for l in open("gpurun_out/r02y_sweep.jsonl"):
    d = json.loads(l); acc[(tuple(d["cells"]), tuple(sorted(d["opts"].items())))].append(d["us_per_iter"])
for k, v in acc.items():
    print(k[0], dict(k[1]), "min %.1f med %.1f" % (min(v), sorted(v)[len(v)//2]), ["%.1f" % x for x in v])
PY
