#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02g_trace.json
for o in "--opt ty=8 --opt kc=64 --opt zshift=1" "--opt ty=8 --opt kc=64 --opt zshift=2" "--opt ty=7 --opt kc=64"; do
  BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 256 $o --out gpurun_out/r02g_trace.json 2>> gpurun_out/r02g_trace.err | cut -c1-300
done
tail -3 gpurun_out/r02g_trace.err
