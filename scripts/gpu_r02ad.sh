#!/bin/bash
# round 2, call AD (1 GPU): register-pipelined epilogue pair as the only streaming form: full epilogue + drop-in tests, smoke, timing per chunk length
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_epilogue.py tests/test_dropin.py -m gpu -x -q ) > gpurun_out/r02ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ad_pytest.log
tail -5 gpurun_out/r02ad_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for o in "epi_chunk=32" "epi_chunk=64" "epi_chunk=128" "epi_chunk=16"; do
  timeout 200 python scripts/epi_profile.py 512 6 $o 2>&1 | tail -1 | cut -c1-160
done | tee gpurun_out/r02ad_epi_times.txt
