#!/bin/bash
# round 2, call Z (8 GPUs, charged 8 x): push-model halo + boundary-first chunks on the full node: 4/8-GPU parity, then 512^3 strong
# scaling with three block shapes (2x2x2 / 1x2x4 / 1x4x2), the channel and the 1000-sphere configuration
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_multiprocess.py -m gpu -x -q -k "four_and_eight" ) > gpurun_out/r02z_pytest_mp_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02z_pytest_mp_n$N.log
tail -6 gpurun_out/r02z_pytest_mp_n$N.log
grep -q "pytest rc=0" gpurun_out/r02z_pytest_mp_n$N.log || exit 0
run() { # tag, extra args...
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-parity "$@" > gpurun_out/r02z_bench_n${N}_$tag.json 2> gpurun_out/r02z_bench_n${N}_$tag.err; echo "bench $tag rc=$?"
  grep '^{' gpurun_out/r02z_bench_n${N}_$tag.json | cut -c1-200; tail -1 gpurun_out/r02z_bench_n${N}_$tag.err
}
run strong512_2x2x2 --blocks 2,2,2
run strong512_1x2x4 --blocks 1,2,4 --no-e2e --no-epilogue
run strong512_1x4x2 --blocks 1,4,2 --no-e2e --no-epilogue
run channel --cells 512,256,256 --bc channel --no-e2e --no-epilogue
run parts1000 --parts 1000 --bc sedimentation --length 64 --no-e2e --no-epilogue
