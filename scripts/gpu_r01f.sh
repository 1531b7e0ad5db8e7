#!/bin/bash
# parity suite (epilogue + face exchanges + PDL policy), golden regeneration, epilogue launch breakdown, bench
set -x
mkdir -p gpurun_out/golden_epi
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python oracle/make_golden_epilogue.py gpurun_out/golden_epi > gpurun_out/golden_epi.log 2>&1; echo "golden rc=$?"; tail -6 gpurun_out/golden_epi.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/epi_launches.csv python scripts/epi_profile.py 512 2 > gpurun_out/epi_profile.log 2>&1; tail -2 gpurun_out/epi_profile.log
grep -E "k_epilogue|k_sub_mean|k_xchg|k_bc_p" gpurun_out/epi_launches.csv | tail -30
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
