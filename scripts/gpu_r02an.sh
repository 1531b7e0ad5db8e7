#!/bin/bash
# round 2, call AN (1 GPU): ncu evidence of the FINAL build: launch list of the bench command + --set full of the two iteration kernels at 512^3
set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02an_launches.csv \
  python bench.py --steps 1 --warmup 1 --fixed-iters 60 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r02an_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02an_prof512 \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02an_ncu512.log 2>&1; tail -2 gpurun_out/r02an_ncu512.log
