#!/bin/bash
mkdir -p gpurun_out
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$3 scripts/comm_probe.py $1 $2 2>gpurun_out/probe_$3.err | grep COMM_PROBE | tee -a gpurun_out/comm_probe.jsonl; tail -2 gpurun_out/probe_$3.err; }
run 1,1,2 512,512,512 1
run 1,1,2 256,256,512 2
run 2,1,1 512,256,256 3
