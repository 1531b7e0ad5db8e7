#!/bin/bash
# round 2, call R (1 GPU): halo of r as a PUSH from the residual kernel (no remote loads in the search kernel): parity
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_cages.py tests/test_dropin.py -m gpu -x -q --durations=5 ) > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -8 gpurun_out/r02r_pytest.log
grep -q "pytest rc=0" gpurun_out/r02r_pytest.log || exit 0
rm -f gpurun_out/r02r_sweep.jsonl
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt pdl=1 --out gpurun_out/r02r_sweep.jsonl > /dev/null 2> gpurun_out/r02r_sweep.err
timeout 300 python scripts/sweep.py --grid 512 --iters 100 --opt pdl=1 --out gpurun_out/r02r_sweep.jsonl > /dev/null 2>> gpurun_out/r02r_sweep.err
cut -c1-330 gpurun_out/r02r_sweep.jsonl; tail -2 gpurun_out/r02r_sweep.err
