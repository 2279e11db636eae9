#!/bin/bash
set -u
O=gpurun_out/r2s8
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_cfg2_n8.json 2> $O/bench_cfg2_n8.err; echo "rc=$?"
