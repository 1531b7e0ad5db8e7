#!/bin/bash
# round 2, call X (1 GPU): final check of the tail changes (L2 scalar loads) + guided-plan sweep at 256^3 and 512x256x128
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py -m gpu -x -q ) > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
tail -5 gpurun_out/r02x_pytest.log
rm -f gpurun_out/r02x_sweep.jsonl
for g in 256,256,256 512,256,128; do
  timeout 300 python scripts/sweep.py --grid $g --iters 200 --opt pdl=1 --opt chunk_min=4,6,8 --opt guided_pct=60,100 --out gpurun_out/r02x_sweep.jsonl > /dev/null 2>> gpurun_out/r02x_sweep.err
done
cut -c1-330 gpurun_out/r02x_sweep.jsonl; tail -2 gpurun_out/r02x_sweep.err
