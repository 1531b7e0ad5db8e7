#!/bin/bash
# round 2, call Q (N GPUs, charged N x): the BASELINE configurations at N ranks on the final tree.
#   N = 8: 512^3 strong (headline), 512x256x256 channel (configs[2]), 1000 spheres 512^3 (configs[3]), 1024^3 (configs[4] strong = 512^3 per GPU weak)
#   N = 4: 512^3 strong, channel;   N = 2: 512^3 strong, channel split in x and in z
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
run() { # tag, extra args...
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-parity "$@" > gpurun_out/r02q_bench_n${N}_$tag.json 2> gpurun_out/r02q_bench_n${N}_$tag.err; echo "bench $tag rc=$?"
  cut -c1-400 gpurun_out/r02q_bench_n${N}_$tag.json; tail -2 gpurun_out/r02q_bench_n${N}_$tag.err
}
run strong512
if [ $N -eq 8 ]; then
  run channel --cells 512,256,256 --bc channel --no-e2e --no-epilogue
  run parts1000 --parts 1000 --bc sedimentation --length 64 --no-e2e --no-epilogue
  run strong1024 --grid 1024 --bc periodic --no-e2e --no-epilogue --steps 2 --warmup 1
elif [ $N -eq 4 ]; then
  run channel --cells 512,256,256 --bc channel --no-e2e --no-epilogue
else
  run channel_x --cells 512,256,256 --bc channel --blocks 2,1,1 --no-e2e --no-epilogue
  run channel_z --cells 512,256,256 --bc channel --blocks 1,1,2 --no-e2e --no-epilogue
fi
