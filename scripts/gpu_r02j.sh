#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02j_trace.json
for o in "" "--opt ty=6 --opt chunk_min=8" "--opt guided=0 --opt ty=6 --opt kc=86"; do
  BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 256 $o --out gpurun_out/r02j_trace.json 2>> gpurun_out/r02j_trace.err | cut -c1-1500
done
tail -3 gpurun_out/r02j_trace.err
