#!/bin/bash
# round 2, call T (N GPUs, charged N x): push-model halo over real NVLink: one-process-per-GPU parity, then the strong-scaling lines
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q --durations=4 ) > gpurun_out/r02t_pytest_mp_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest_mp_n$N.log
tail -10 gpurun_out/r02t_pytest_mp_n$N.log
grep -q "pytest rc=0" gpurun_out/r02t_pytest_mp_n$N.log || exit 0
run() { # tag, extra args...
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-parity "$@" > gpurun_out/r02t_bench_n${N}_$tag.json 2> gpurun_out/r02t_bench_n${N}_$tag.err; echo "bench $tag rc=$?"
  cut -c1-300 gpurun_out/r02t_bench_n${N}_$tag.json; tail -2 gpurun_out/r02t_bench_n${N}_$tag.err
}
run strong512
if [ $N -eq 8 ]; then
  run channel --cells 512,256,256 --bc channel --no-e2e --no-epilogue
  run parts1000 --parts 1000 --bc sedimentation --length 64 --no-e2e --no-epilogue
elif [ $N -eq 2 ]; then
  run channel_x --cells 512,256,256 --bc channel --blocks 2,1,1 --no-e2e --no-epilogue
fi
