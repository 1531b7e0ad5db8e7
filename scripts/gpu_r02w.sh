#!/bin/bash
# round 2, call W (1 GPU): is the comm time-out of test_full_and_ragged_tiles[cells4-blocks4-False-5-11] reproducible?
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "test_full_and_ragged_tiles" 2>&1 | tail -4
done > gpurun_out/r02w_ragged.log 2>&1
cat gpurun_out/r02w_ragged.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_dropin.py tests/test_gpu_epilogue.py -m gpu -q ) > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_pytest.log
tail -12 gpurun_out/r02w_pytest.log
