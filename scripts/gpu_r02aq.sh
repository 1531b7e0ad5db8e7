#!/bin/bash
# round 2, call AQ (1 GPU): the planner as an exported host function (make_plan calls it): smoke + the plan-dependent parity tests
mkdir -p gpurun_out
( timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan or tile_height or ragged or full_and" 2>&1 | tail -2 ) > gpurun_out/r02aq.log 2>&1
cat gpurun_out/r02aq.log
