#!/bin/bash
# compute-sanitizer passes: memcheck, synccheck, initcheck on every kernel of the library (1 rank and 2 x 1 x 1
# ranks in one process), racecheck on a tiny solve.  Logs under gpurun_out/<tag>_sanitizer_*.log; usage: gpu_sanitize.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck synccheck initcheck; do
  for ranks in 1; do        # 2 single-process ranks cannot run under the tool (it serialises the streams whose kernels wait for each other): profiles/r02c_sanitizer.md
    log=gpurun_out/${tag}_sanitizer_${tool}_n${ranks}.log
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 python scripts/sanitize_case.py --ranks $ranks > $log 2>&1
    echo "exit code $?" >> $log
    grep -E "SANITIZE_CASE_OK|ERROR SUMMARY|exit code|solve " $log | tail -8
  done
done
log=gpurun_out/${tag}_sanitizer_racecheck_n1.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 20 python scripts/sanitize_case.py --small > $log 2>&1
echo "exit code $?" >> $log
grep -E "SANITIZE_CASE_OK|RACECHECK SUMMARY|exit code|solve " $log | tail -6
