#!/bin/bash
# multi-GPU pass: real one-process-per-GPU parity + the strong-scaling bench line at N = number of visible GPUs
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_multiprocess.py -m gpu -x -q > gpurun_out/pytest_mp_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mp_n$N.log
tail -30 gpurun_out/pytest_mp_n$N.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
    tail -1 gpurun_out/bench_n$n.json
    tail -5 gpurun_out/bench_n$n.err
  fi
done
