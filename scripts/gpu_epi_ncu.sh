#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_epilogue|k_sub_mean' -c 2 -f -o gpurun_out/prof_epi \
  python scripts/epi_profile.py 512 1 > gpurun_out/ncu_epi.log 2>&1
tail -3 gpurun_out/ncu_epi.log; ls -la gpurun_out/prof_epi.ncu-rep
