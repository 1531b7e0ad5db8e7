#!/bin/bash
# round 2, call D (1 GPU): select-form operator + producer-warp TMA issue: parity, BC_star / prologue / timed drop-in tests,
# golden vectors of the reference's BC_* kernels, A/B sweeps of the producer warp at 256^3 and 512^3, particle bench.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_bc_star.py tests/test_dropin.py -m gpu -x -q --durations=5 ) > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
timeout 300 python oracle/make_golden_bcstar.py gpurun_out/golden_bcs > gpurun_out/r02d_golden_bcs.log 2>&1; tail -4 gpurun_out/r02d_golden_bcs.log
rm -f gpurun_out/r02d_sweep256.jsonl gpurun_out/r02d_sweep512.jsonl
timeout 400 python scripts/sweep.py --grid 256 --iters 200 --opt tma_warp=0,1 --opt ty=8,7,6,5,4 --opt kc=128,86,64,43,32,22,16 --opt pdl=1 --out gpurun_out/r02d_sweep256.jsonl > /dev/null 2> gpurun_out/r02d_sweep.err
cut -c1-330 gpurun_out/r02d_sweep256.jsonl
timeout 300 python scripts/sweep.py --grid 512 --iters 100 --opt tma_warp=0,1 --opt ty=8 --opt kc=24,32 --opt pdl=1 --out gpurun_out/r02d_sweep512.jsonl > /dev/null 2>> gpurun_out/r02d_sweep.err
cut -c1-330 gpurun_out/r02d_sweep512.jsonl
tail -3 gpurun_out/r02d_sweep.err
timeout 600 python bench.py --parts 1000 --bc sedimentation --length 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-epilogue > gpurun_out/r02d_bench_parts1000.json 2> gpurun_out/r02d_bench_parts1000.err; cut -c1-2500 gpurun_out/r02d_bench_parts1000.json; tail -3 gpurun_out/r02d_bench_parts1000.err
