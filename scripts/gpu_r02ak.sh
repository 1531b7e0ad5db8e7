#!/bin/bash
# round 2, call AK (1 GPU): every push of the residual kernel (x faces too) behind the call: parity, instruction count of the kernel, timing
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -m gpu -x -q ) > gpurun_out/r02ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ak_pytest.log
tail -4 gpurun_out/r02ak_pytest.log
rm -f gpurun_out/r02ak_sweep.jsonl
timeout 300 python scripts/sweep.py --grid 512 --iters 100 --repeat 2 --opt pdl=1 --out gpurun_out/r02ak_sweep.jsonl > /dev/null 2> gpurun_out/r02ak_sweep.err
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --repeat 2 --opt pdl=1 --out gpurun_out/r02ak_sweep.jsonl > /dev/null 2>> gpurun_out/r02ak_sweep.err
cut -c1-330 gpurun_out/r02ak_sweep.jsonl; tail -2 gpurun_out/r02ak_sweep.err
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'k_search_tma|k_resid_tma' -s 20 -c 4 --csv --log-file gpurun_out/r02ak_inst.csv \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02ak_ncu.log 2>&1
grep -E "k_(search|resid)_tma" gpurun_out/r02ak_inst.csv | cut -d, -f5,13- | head -8
