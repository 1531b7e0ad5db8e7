#!/bin/bash
# GPU pass for the solve epilogue: parity tests, golden vectors from the reference's own kernels, smoke, both bench arms
set -x
mkdir -p gpurun_out/golden_epi
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python oracle/make_golden_epilogue.py gpurun_out/golden_epi > gpurun_out/golden_epi.log 2>&1; echo "golden rc=$?"; cat gpurun_out/golden_epi.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
timeout 500 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref_n1.err
cat gpurun_out/bench_ref_n1.json
