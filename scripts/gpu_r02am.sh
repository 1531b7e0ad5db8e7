#!/bin/bash
# round 2, call AM (1 GPU): the full GPU suite and the default bench on the final tree
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r02am_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02am_pytest_gpu.log
tail -12 gpurun_out/r02am_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02am_bench_reference_n1.json 2> gpurun_out/r02am_bench_reference_n1.err; cut -c1-200 gpurun_out/r02am_bench_reference_n1.json
timeout 600 python bench.py > gpurun_out/r02am_bench_n1.json 2> gpurun_out/r02am_bench_n1.err; cut -c1-400 gpurun_out/r02am_bench_n1.json; tail -2 gpurun_out/r02am_bench_n1.err
