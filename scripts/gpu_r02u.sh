#!/bin/bash
# round 2, call U (1 GPU): tail split of the iteration kernels at 256^3 (trace build); stand-alone time of the candidate per-rank blocks of an 8-GPU run
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02u_trace.json gpurun_out/r02u_shapes.jsonl
BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 256 --out gpurun_out/r02u_trace.json 2> gpurun_out/r02u_trace.err | cut -c1-2500
for g in 256,256,256 512,256,128 256,512,128 512,512,64 512,128,256 128,512,256; do
  timeout 200 python scripts/sweep.py --grid $g --iters 200 --opt pdl=1 --out gpurun_out/r02u_shapes.jsonl > /dev/null 2>> gpurun_out/r02u_shapes.err
done
cut -c1-330 gpurun_out/r02u_shapes.jsonl; tail -2 gpurun_out/r02u_shapes.err
