#!/bin/bash
# lean multi-GPU pass (8-GPU box time is charged 8x): one real-NVLink parity run at 8 ranks, then strong-scaling bench lines
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
run() { # n blocks tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 \
    bench.py --gpus $1 --steps 3 --warmup 3 --blocks $2 > gpurun_out/bench_n$1_$3.json 2> gpurun_out/bench_n$1_$3.err; echo "bench n=$1 blocks=$2 rc=$?"
}
if [ $N -ge 8 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 \
    tests/mp_worker.py --mode gpu --cells 48,40,44 --blocks 2,2,2 --bc periodic 2>&1 | grep MP_WORKER_OK
  run 8 2,2,2 222; run 8 1,2,4 124; run 4 1,2,2 122; run 4 1,1,4 114
elif [ $N -ge 2 ]; then
  run 2 1,1,2 112
fi
