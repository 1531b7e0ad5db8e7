#!/bin/bash
# round 2, call K (1 GPU): planner inputs for the guided regime at the per-rank shapes of the 2/4/8-GPU runs; cost split of the particle kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02k_sweep*.jsonl
for g in 256,256,256 512,256,256 256,128,128 512,512,256; do
  timeout 300 python scripts/sweep.py --grid $g --iters 150 --opt ty=8,7,6 --opt chunk_min=8,12 --opt pdl=1 --out gpurun_out/r02k_sweep_shapes.jsonl > /dev/null 2>> gpurun_out/r02k_sweep.err
done
cut -c1-330 gpurun_out/r02k_sweep_shapes.jsonl
tail -3 gpurun_out/r02k_sweep.err
for p in 0 1 1000; do
  timeout 400 python bench.py --parts $p --bc sedimentation --length 64 --steps 2 --warmup 1 --fixed-iters 100 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02k_bench_parts$p.json 2> gpurun_out/r02k_bench_parts$p.err; cut -c1-200 gpurun_out/r02k_bench_parts$p.json; tail -2 gpurun_out/r02k_bench_parts$p.err
done
