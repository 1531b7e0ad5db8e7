#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29588 \
  scripts/decomp_probe.py "$1" 512 duct 2> gpurun_out/probe8.err | grep DECOMP_PROBE | tee -a gpurun_out/decomp_probe.jsonl
tail -3 gpurun_out/probe8.err
