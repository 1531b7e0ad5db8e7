#!/bin/bash
# round 2, call E (1 GPU): coefficient producers (cages) parity + golden vectors; per-CTA time line of the iteration kernels at
# the 8-GPU block size (debug build); ncu source-level capture of both kernels at 256^3.
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_cages.py tests/test_abi.py -m gpu -x -q --durations=5 ) > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -15 gpurun_out/r02e_pytest.log
timeout 300 python oracle/make_golden_cages.py gpurun_out/golden_cages > gpurun_out/r02e_golden_cages.log 2>&1; tail -6 gpurun_out/r02e_golden_cages.log
rm -f gpurun_out/r02e_trace.json
for o in "--opt ty=8 --opt kc=64" "--opt ty=6 --opt kc=86" "--opt ty=8 --opt kc=128" "--opt ty=8 --opt kc=32" "--opt ty=8 --opt kc=64 --opt pdl=0"; do
  BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 256 $o --out gpurun_out/r02e_trace.json 2>> gpurun_out/r02e_trace.err | cut -c1-1800
done
BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 512 --iters 10 --out gpurun_out/r02e_trace.json 2>> gpurun_out/r02e_trace.err | cut -c1-1800
tail -3 gpurun_out/r02e_trace.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_search_tma|k_resid_tma' -s 20 -c 2 -f -o gpurun_out/r02e_prof256 \
  python bench.py --grid 256 --steps 1 --warmup 0 --fixed-iters 30 --no-cpu-baseline --no-e2e --no-epilogue > gpurun_out/r02e_ncu_full.log 2>&1
tail -3 gpurun_out/r02e_ncu_full.log
ls -la gpurun_out | tail -12
