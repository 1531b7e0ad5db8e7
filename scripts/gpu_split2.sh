#!/bin/bash
# N = 2: exposed halo + all-reduce time for the three split directions (x: element-strided faces, y: rows, z: planes)
for b in 2,1,1 1,2,1 1,1,2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 2 --warmup 1 --blocks $b --no-e2e --grid ${GRID:-512} > gpurun_out/bench_split_$b.json 2> gpurun_out/bench_split_$b.err
  tail -1 gpurun_out/bench_split_$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$b', 'value %.1f'%d['value'], {k:round(v,1) for k,v in d['comm'].items() if k!='method'})"
done
