#!/bin/bash
# round 2, call P (1 GPU): evidence on the final single-GPU tree: full GPU suite, smoke, both bench arms at 512^3 (parity object),
# BASELINE configs 1 (256^3), 4 at N = 1 (1024^3), ncu launch list + --set full captures of the iteration kernels at 512^3 and 256^3
# and of the particle instantiations.
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 ) > gpurun_out/r02p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest_gpu.log
tail -18 gpurun_out/r02p_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02p_smoke.log 2>&1; tail -3 gpurun_out/r02p_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02p_bench_reference_n1.json 2> gpurun_out/r02p_bench_reference_n1.err; cut -c1-300 gpurun_out/r02p_bench_reference_n1.json
timeout 600 python bench.py > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; cut -c1-1200 gpurun_out/r02p_bench_n1.json; tail -3 gpurun_out/r02p_bench_n1.err
timeout 600 python bench.py --impl reference --grid 256 --steps 3 --warmup 1 > gpurun_out/r02p_bench_reference_256.json 2> gpurun_out/r02p_bench_reference_256.err
timeout 400 python bench.py --grid 256 --no-cpu-baseline > gpurun_out/r02p_bench_256.json 2> gpurun_out/r02p_bench_256.err; cut -c1-300 gpurun_out/r02p_bench_256.json; tail -2 gpurun_out/r02p_bench_256.err
timeout 900 python bench.py --grid 1024 --bc periodic --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-epilogue > gpurun_out/r02p_bench_1024.json 2> gpurun_out/r02p_bench_1024.err; cut -c1-300 gpurun_out/r02p_bench_1024.json; tail -3 gpurun_out/r02p_bench_1024.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p_launches.csv \
  python bench.py --steps 1 --warmup 1 --fixed-iters 60 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r02p_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02p_prof512 \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02p_ncu512.log 2>&1; tail -2 gpurun_out/r02p_ncu512.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02p_prof256 \
  python bench.py --grid 256 --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02p_ncu256.log 2>&1; tail -2 gpurun_out/r02p_ncu256.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02p_prof_parts \
  python bench.py --grid 256 --parts 125 --bc sedimentation --length 32 --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02p_ncu_parts.log 2>&1; tail -2 gpurun_out/r02p_ncu_parts.log
ls -la gpurun_out | tail -20
