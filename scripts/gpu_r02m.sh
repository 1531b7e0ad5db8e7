#!/bin/bash
# round 2, call M (1 GPU): particle factors from the solid bit of the mask tile (no pmask stream) + streaming epilogue pair: parity, then benches
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_bench_configs.py tests/test_cages.py tests/test_gpu_epilogue.py -m gpu -x -q --durations=5 ) > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
tail -12 gpurun_out/r02m_pytest.log
grep -q "pytest rc=0" gpurun_out/r02m_pytest.log || exit 0
for p in 1 1000; do
  timeout 400 python bench.py --parts $p --bc sedimentation --length 64 --steps 2 --warmup 1 --fixed-iters 100 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r02m_bench_parts$p.json 2> gpurun_out/r02m_bench_parts$p.err; cut -c1-200 gpurun_out/r02m_bench_parts$p.json; tail -2 gpurun_out/r02m_bench_parts$p.err
done
for o in "epilogue_tiled=1" "epi_chunk=16" "epi_chunk=32" "epi_chunk=64" "epi_chunk=128"; do
  timeout 300 python bench.py --steps 1 --warmup 0 --fixed-iters 20 --no-cpu-baseline --no-e2e --no-parity --opt $o > gpurun_out/r02m_epi_$o.json 2> gpurun_out/r02m_epi.err; python -c "
import json,sys; j=json.loads(open('gpurun_out/r02m_epi_$o.json').read().strip().splitlines()[-1]); print('$o', j['epilogue']['ms_per_call'], j['epilogue']['frac'])"
done
tail -2 gpurun_out/r02m_epi.err
