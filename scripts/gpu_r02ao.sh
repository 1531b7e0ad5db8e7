#!/bin/bash
# round 2, call AO (1 GPU): per-kernel events on a 64-iteration sample of the first timed step; the timed entry point still times every iteration
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_dropin.py tests/test_gpu_parity.py -m gpu -x -q -k "timed or kernel_timing or noparts_all_bc" ) > gpurun_out/r02ao_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ao_pytest.log; tail -3 gpurun_out/r02ao_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/r02ao_bench_n1.json 2> gpurun_out/r02ao_bench_n1.err; cut -c1-300 gpurun_out/r02ao_bench_n1.json; tail -2 gpurun_out/r02ao_bench_n1.err
