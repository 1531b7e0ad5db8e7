#!/bin/bash
# session-4 GPU pass on HEAD: parity tests, bench, ncu launch list, full capture of the two iteration kernels
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/nvidia_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 4 -f -o gpurun_out/prof_iter \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
