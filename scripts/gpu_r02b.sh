#!/bin/bash
# round 2, call B (1 GPU): first run of the reworked iteration kernels (run-time tile height, XFULL / all-ones-mask fast paths,
# slot-filling planner): parity, then option sweeps at the 8-GPU block size and at 512^3, then ncu of both kernels.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q --durations=8 ) > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -25 gpurun_out/r02b_pytest.log
rm -f gpurun_out/r02b_sweep256.jsonl gpurun_out/r02b_sweep512.jsonl
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt ty=0,8,7,6,5 --opt kc=0 --opt pdl=0,1 --out gpurun_out/r02b_sweep256.jsonl > /dev/null 2> gpurun_out/r02b_sweep.err
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt ty=8,7 --opt kc=32,128 --opt pdl=0 --out gpurun_out/r02b_sweep256.jsonl > /dev/null 2>> gpurun_out/r02b_sweep.err
cut -c1-330 gpurun_out/r02b_sweep256.jsonl
timeout 400 python scripts/sweep.py --grid 512 --iters 100 --opt ty=0,8,7,6 --opt kc=0,24,128 --opt pdl=1 --out gpurun_out/r02b_sweep512.jsonl > /dev/null 2>> gpurun_out/r02b_sweep.err
cut -c1-330 gpurun_out/r02b_sweep512.jsonl
tail -5 gpurun_out/r02b_sweep.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02b_prof512 \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue > gpurun_out/r02b_ncu_full.log 2>&1
tail -3 gpurun_out/r02b_ncu_full.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-epilogue > gpurun_out/r02b_bench512.json 2> gpurun_out/r02b_bench512.err; cut -c1-1200 gpurun_out/r02b_bench512.json; tail -3 gpurun_out/r02b_bench512.err
