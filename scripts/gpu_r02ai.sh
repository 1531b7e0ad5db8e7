#!/bin/bash
# round 2, call AI (1 GPU): per-plane gating of the y / z push calls: parity of the decomposed solves
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_configs.py tests/test_gpu_reference.py -m gpu -x -q ) > gpurun_out/r02ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ai_pytest.log
tail -4 gpurun_out/r02ai_pytest.log
