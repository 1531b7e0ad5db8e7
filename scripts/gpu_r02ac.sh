#!/bin/bash
# round 2, call AC (1 GPU): hybrid k_epi_uwp_reg (registers for the i-lane streams, cp.async for u*, 3 CTAs / SM): parity, timing, ncu
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_epilogue.py -m gpu -x -q -k "three_way or streaming_pair" ) > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
tail -4 gpurun_out/r02ac_pytest.log
for o in "epi_v_reg=0 epi_uwp_reg=0" "epi_v_reg=1 epi_uwp_reg=1" "epi_v_reg=1 epi_uwp_reg=1 epi_chunk=64" "epi_v_reg=1 epi_uwp_reg=1 epi_chunk=16"; do
  timeout 200 python scripts/epi_profile.py 512 6 $o 2>&1 | tail -1 | cut -c1-160
done | tee gpurun_out/r02ac_epi_times.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_epi_uwp|k_epi_v|k_sub_mean' -c 3 -f -o gpurun_out/r02ac_epi_prof python scripts/epi_profile.py 512 1 epi_v_reg=1 epi_uwp_reg=1 > gpurun_out/r02ac_ncu.log 2>&1; tail -2 gpurun_out/r02ac_ncu.log
