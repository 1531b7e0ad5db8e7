#!/bin/bash
# round 2, call N (1 GPU): FM_NEAR encoding (particle gathers only near particles): parity, then the particle bench
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_cages.py -m gpu -x -q --durations=5 ) > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
tail -8 gpurun_out/r02n_pytest.log
grep -q "pytest rc=0" gpurun_out/r02n_pytest.log || exit 0
for p in 0 1 1000; do
  timeout 400 python bench.py --parts $p --bc sedimentation --length 64 --steps 2 --warmup 1 --fixed-iters 100 --no-cpu-baseline --no-e2e --no-parity --no-epilogue > gpurun_out/r02n_bench_parts$p.json 2> gpurun_out/r02n_bench_parts$p.err; cut -c1-200 gpurun_out/r02n_bench_parts$p.json; tail -2 gpurun_out/r02n_bench_parts$p.err
done
