#!/bin/bash
# 2-GPU pass (charged 2x): real one-process-per-GPU parity over NVLink (incl. refresh path and epilogue), then N=2 bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q > gpurun_out/pytest_mp_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n2.log
tail -15 gpurun_out/pytest_mp_n2.log
run() { # blocks tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$2 \
    bench.py --gpus 2 --steps 3 --warmup 3 --blocks $1 --no-e2e > gpurun_out/bench_n2_$2.json 2> gpurun_out/bench_n2_$2.err; echo "bench n=2 blocks=$1 rc=$?"
  tail -1 gpurun_out/bench_n2_$2.json; tail -3 gpurun_out/bench_n2_$2.err
}
run 1,1,2 1; run 2,1,1 2
