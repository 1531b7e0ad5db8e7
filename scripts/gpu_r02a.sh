#!/bin/bash
# round 2, call A (1 GPU): small-block (256^3 = the 8-GPU share of 512^3) evidence before touching the kernels:
# option sweep with per-kernel events, ncu launch list and --set full of both iteration kernels at 256^3, sustained HBM copy.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/nvidia_smi.txt 2>&1
timeout 120 python scripts/hbm_sustained.py > gpurun_out/r02a_hbm.json 2> gpurun_out/r02a_hbm.err; cat gpurun_out/r02a_hbm.json
rm -f gpurun_out/r02a_sweep256.jsonl
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt tile=0,1,8 --opt kc=0,22,29,32,43,64 --opt pdl=0,1 --out gpurun_out/r02a_sweep256.jsonl > /dev/null 2> gpurun_out/r02a_sweep.err
cat gpurun_out/r02a_sweep256.jsonl | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches256.csv \
  python bench.py --grid 256 --steps 1 --warmup 1 --fixed-iters 60 --no-cpu-baseline --no-e2e --no-epilogue > gpurun_out/r02a_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 4 -f -o gpurun_out/r02a_prof256 \
  python bench.py --grid 256 --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue > gpurun_out/r02a_ncu_full.log 2>&1
tail -3 gpurun_out/r02a_ncu_full.log
timeout 300 python bench.py --grid 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench256.json 2> gpurun_out/r02a_bench256.err; cat gpurun_out/r02a_bench256.json | cut -c1-1500
ls -la gpurun_out
