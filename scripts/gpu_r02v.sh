#!/bin/bash
# round 2, call V (1 GPU): batched item-order sum, preloaded tail scalars, boundary-first z-chunks: parity + sweeps + tail split
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_dropin.py -m gpu -x -q ) > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log
tail -6 gpurun_out/r02v_pytest.log
rm -f gpurun_out/r02v_trace.json gpurun_out/r02v_sweep.jsonl
BBPCG_LIB_PATH=$PWD/bluebottle-3.0_b200/lib/libbbpcg_trace.so timeout 200 python scripts/trace_timeline.py --grid 256 --out gpurun_out/r02v_trace.json 2> gpurun_out/r02v_trace.err | cut -c1-1800
for g in 256,256,256 512,256,128 512,128,256 512,512,512; do
  timeout 200 python scripts/sweep.py --grid $g --iters 200 --opt pdl=1 --out gpurun_out/r02v_sweep.jsonl > /dev/null 2>> gpurun_out/r02v_sweep.err
done
cut -c1-330 gpurun_out/r02v_sweep.jsonl; tail -2 gpurun_out/r02v_sweep.err
