#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02s_sweep.jsonl
timeout 300 python scripts/sweep.py --grid 256 --iters 200 --opt pdl=1 --out gpurun_out/r02s_sweep.jsonl > /dev/null 2> gpurun_out/r02s_sweep.err
timeout 300 python scripts/sweep.py --grid 512 --iters 100 --opt pdl=1 --out gpurun_out/r02s_sweep.jsonl > /dev/null 2>> gpurun_out/r02s_sweep.err
timeout 300 python scripts/sweep.py --grid 512 --bc cavity --iters 100 --opt pdl=1 --out gpurun_out/r02s_sweep.jsonl > /dev/null 2>> gpurun_out/r02s_sweep.err
cut -c1-330 gpurun_out/r02s_sweep.jsonl; tail -2 gpurun_out/r02s_sweep.err
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r02s_pytest.log 2>&1; tail -3 gpurun_out/r02s_pytest.log
