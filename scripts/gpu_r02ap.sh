#!/bin/bash
# round 2, call AP (N GPUs, charged N x): the default strong-scaling line of the final tree at N ranks (the driver's SCALE command)
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --no-cpu-baseline --no-parity > gpurun_out/r02ap_bench_n${N}.json 2> gpurun_out/r02ap_bench_n${N}.err; echo "bench rc=$?"
grep '^{' gpurun_out/r02ap_bench_n${N}.json | cut -c1-250; tail -2 gpurun_out/r02ap_bench_n${N}.err
