#!/bin/bash
# round 2, call AE (1 GPU): evidence on the final tree: full GPU suite, smoke, both bench arms at 512^3 (parity object), 256^3,
# ncu launch list + --set full captures of the iteration kernels at 512^3 and 256^3
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02ae_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ae_pytest_gpu.log
tail -14 gpurun_out/r02ae_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02ae_smoke.log 2>&1; tail -3 gpurun_out/r02ae_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02ae_bench_reference_n1.json 2> gpurun_out/r02ae_bench_reference_n1.err; cut -c1-300 gpurun_out/r02ae_bench_reference_n1.json
timeout 600 python bench.py > gpurun_out/r02ae_bench_n1.json 2> gpurun_out/r02ae_bench_n1.err; cut -c1-1500 gpurun_out/r02ae_bench_n1.json; tail -3 gpurun_out/r02ae_bench_n1.err
timeout 400 python bench.py --grid 256 --no-cpu-baseline --no-parity > gpurun_out/r02ae_bench_256.json 2> gpurun_out/r02ae_bench_256.err; cut -c1-300 gpurun_out/r02ae_bench_256.json; tail -2 gpurun_out/r02ae_bench_256.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02ae_launches.csv \
  python bench.py --steps 1 --warmup 1 --fixed-iters 60 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r02ae_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02ae_prof512 \
  python bench.py --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02ae_ncu512.log 2>&1; tail -2 gpurun_out/r02ae_ncu512.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02ae_prof256 \
  python bench.py --grid 256 --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue --no-parity > gpurun_out/r02ae_ncu256.log 2>&1; tail -2 gpurun_out/r02ae_ncu256.log
ls -la gpurun_out | tail -12
