#!/bin/bash
# 8-GPU call: cfg2 weak scaling point, cfg3 (2 M Gaussians, 8 views, one per GPU), cfg4 full loop on 8 GPUs
set -u
O=gpurun_out/r2s8
mkdir -p $O
N=${1:-8}
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run bench_cfg2_n$N --steps 20 --warmup 5
run bench_cfg3_n$N --config cfg3 --steps 20 --warmup 5
run loop_cfg4_n$N --config cfg4
nvidia-smi topo -m > $O/topo.txt 2>&1
grep -h "Error\|error" $O/*.err | head -5
