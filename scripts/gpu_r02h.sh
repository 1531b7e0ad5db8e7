#!/bin/bash
# round 2, call H (8 GPUs, charged 8x: keep it short): first-ever 4- and 8-GPU one-process-per-GPU parity run, then the
# strong-scaling line at N = 8 with the default plan and with ty 6 / kc 86, and BASELINE config 3 (1000 spheres, 512^3 / 8).
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02h_topo_n$N.txt 2>&1
( time timeout 600 python -m pytest tests/test_multiprocess.py -m gpu -x -q --durations=6 ) > gpurun_out/r02h_pytest_mp_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest_mp_n$N.log
tail -14 gpurun_out/r02h_pytest_mp_n$N.log
run() { # tag, extra args...
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 \
    bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-parity "$@" > gpurun_out/r02h_bench_n${N}_$tag.json 2> gpurun_out/r02h_bench_n${N}_$tag.err; echo "bench $tag rc=$?"
  cut -c1-900 gpurun_out/r02h_bench_n${N}_$tag.json; tail -2 gpurun_out/r02h_bench_n${N}_$tag.err
}
run default
run ty6 --ty 6 --kc 86
run parts1000 --parts 1000 --bc sedimentation --length 64 --no-e2e --no-epilogue
