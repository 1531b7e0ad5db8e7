#!/bin/bash
# round 2, call AB (1 GPU): register-pipelined epilogue pair with ld.global.nc instead of evict-first loads (which lost the L2 hits on shared sectors)
set -x
mkdir -p gpurun_out
for o in "epi_v_reg=0 epi_uwp_reg=0" "epi_v_reg=1 epi_uwp_reg=0" "epi_v_reg=0 epi_uwp_reg=1" "epi_v_reg=1 epi_uwp_reg=1"; do
  timeout 200 python scripts/epi_profile.py 512 6 $o 2>&1 | tail -1 | cut -c1-140
done | tee gpurun_out/r02ab_epi_times.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_epi_uwp|k_epi_v' -c 2 -f -o gpurun_out/r02ab_epi_prof python scripts/epi_profile.py 512 1 epi_v_reg=1 epi_uwp_reg=1 > gpurun_out/r02ab_ncu.log 2>&1; tail -2 gpurun_out/r02ab_ncu.log
