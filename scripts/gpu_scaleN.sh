#!/bin/bash
# cfg2 weak-scaling point at N GPUs
set -u
N=${1:-2}
O=gpurun_out/r2s$N
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_cfg2_n$N.json 2> $O/bench_cfg2_n$N.err; echo "rc=$?"
