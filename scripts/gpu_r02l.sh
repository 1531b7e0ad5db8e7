#!/bin/bash
# round 2, call L (2 GPUs): exposed communication of the persistent kernels with 256^3 per rank, PDL on / off
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02l_comm_probe.jsonl
for cfg in "1,1,2 256,256,512" "2,1,1 512,256,256" "1,2,1 256,512,256"; do
  set -- $cfg
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/comm_probe.py $1 $2 2> gpurun_out/r02l_probe.err | grep COMM_PROBE | sed 's/COMM_PROBE //' >> gpurun_out/r02l_comm_probe.jsonl
done
cat gpurun_out/r02l_comm_probe.jsonl
tail -3 gpurun_out/r02l_probe.err
