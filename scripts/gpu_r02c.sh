#!/bin/bash
# round 2, call C (1 GPU): full GPU suite on the reworked kernels (incl. the benchmarked-size parity tests against the
# reference's own kernels), smoke, both bench arms at 512^3 (parity object), 256^3, the 1000-sphere case, sanitizers.
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_gpu.log
tail -22 gpurun_out/r02c_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02c_smoke.log 2>&1; tail -3 gpurun_out/r02c_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_reference_n1.json 2> gpurun_out/r02c_bench_reference_n1.err; cut -c1-600 gpurun_out/r02c_bench_reference_n1.json
timeout 600 python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; cut -c1-3000 gpurun_out/r02c_bench_n1.json; tail -3 gpurun_out/r02c_bench_n1.err
timeout 400 python bench.py --grid 256 > gpurun_out/r02c_bench_256.json 2> gpurun_out/r02c_bench_256.err; cut -c1-400 gpurun_out/r02c_bench_256.json; tail -3 gpurun_out/r02c_bench_256.err
timeout 900 python bench.py --parts 1000 --bc sedimentation --length 64 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02c_bench_parts1000.json 2> gpurun_out/r02c_bench_parts1000.err; cut -c1-3000 gpurun_out/r02c_bench_parts1000.json; tail -5 gpurun_out/r02c_bench_parts1000.err
bash scripts/gpu_sanitize.sh r02c
ls -la gpurun_out | tail -20
