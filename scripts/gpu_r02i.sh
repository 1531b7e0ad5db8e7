#!/bin/bash
# round 2, call I (1 GPU): persistent item-claiming iteration kernels: parity first (under a hard timeout), then sweeps.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_cages.py -m gpu -x -q --durations=5 ) > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -15 gpurun_out/r02i_pytest.log
grep -q "pytest rc=0" gpurun_out/r02i_pytest.log || exit 0
rm -f gpurun_out/r02i_sweep256.jsonl gpurun_out/r02i_sweep512.jsonl
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt guided=1 --opt ty=8,6 --opt chunk_min=8,12,16,24 --opt guided_pct=60,100,150 --opt pdl=1 --out gpurun_out/r02i_sweep256.jsonl > /dev/null 2> gpurun_out/r02i_sweep.err
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt guided=0 --opt ty=8,6 --opt kc=16,24,32,43,64,86 --opt pdl=1 --out gpurun_out/r02i_sweep256.jsonl > /dev/null 2>> gpurun_out/r02i_sweep.err
cut -c1-360 gpurun_out/r02i_sweep256.jsonl
timeout 300 python scripts/sweep.py --grid 512 --iters 100 --opt ty=8 --opt kc=0,24,32 --opt pdl=1 --out gpurun_out/r02i_sweep512.jsonl > /dev/null 2>> gpurun_out/r02i_sweep.err
cut -c1-360 gpurun_out/r02i_sweep512.jsonl
tail -3 gpurun_out/r02i_sweep.err
