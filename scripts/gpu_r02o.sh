#!/bin/bash
# round 2, call O (1 GPU): cp.async-pipelined epilogue pair: parity, timing per chunk length, ncu of both kernels
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_epilogue.py tests/test_dropin.py -m gpu -x -q --durations=5 ) > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -8 gpurun_out/r02o_pytest.log
grep -q "pytest rc=0" gpurun_out/r02o_pytest.log || exit 0
for o in "epilogue_tiled=1" "epi_chunk=16" "epi_chunk=32" "epi_chunk=64"; do
  timeout 300 python bench.py --steps 1 --warmup 0 --fixed-iters 20 --no-cpu-baseline --no-e2e --no-parity --opt $o > gpurun_out/r02o_epi_$o.json 2> gpurun_out/r02o_epi.err; python -c "
import json,sys; j=json.loads(open('gpurun_out/r02o_epi_$o.json').read().strip().splitlines()[-1]); print('$o', j['epilogue']['ms_per_call'], j['epilogue']['frac'])"
done
tail -2 gpurun_out/r02o_epi.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_epi_uwp|k_epi_v|k_sub_mean' -c 3 -f -o gpurun_out/r02o_epi_prof python scripts/epi_profile.py > gpurun_out/r02o_ncu.log 2>&1; tail -3 gpurun_out/r02o_ncu.log
