#!/bin/bash
# round 2, call AR (1 GPU): last look at the in-tree library: smoke()
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/r02ar_smoke.log; cat gpurun_out/r02ar_smoke.log
