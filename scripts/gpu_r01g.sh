#!/bin/bash
# narrow pass: the tests added since r01f (solvability, restart replay) + golden regeneration with the solvability vectors
set -x
mkdir -p gpurun_out/golden_epi
( time timeout 600 python -m pytest tests/test_gpu_epilogue.py tests/test_dropin.py -m gpu -x -q ) > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
tail -25 gpurun_out/pytest_gpu_g.log
timeout 300 python oracle/make_golden_epilogue.py gpurun_out/golden_epi > gpurun_out/golden_epi.log 2>&1; echo "golden rc=$?"; tail -6 gpurun_out/golden_epi.log
