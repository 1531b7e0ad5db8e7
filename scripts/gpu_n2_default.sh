#!/bin/bash
# default bench.py path at N=2 (what the driver's scaling run launches): solve + comm split + e2e + epilogue + solvability over NVLink
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_default.json 2> gpurun_out/bench_n2_default.err; echo "rc=$?"
tail -1 gpurun_out/bench_n2_default.json | cut -c1-3000; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_n2_default.err | tail -5
