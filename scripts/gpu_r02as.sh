#!/bin/bash
# round 2, call AS (1 GPU): 256^3 (BASELINE configs[1]) on the final build
mkdir -p gpurun_out
timeout 90 python bench.py --grid 256 --no-cpu-baseline --no-parity > gpurun_out/r02as_bench_256.json 2> gpurun_out/r02as_bench_256.err; cut -c1-200 gpurun_out/r02as_bench_256.json
