#!/bin/bash
# 2-GPU call: bench at N = 2 (mapper exchange inside value; C++ host legs) + the cfg2 loop on 2 GPUs.
set -u
O=gpurun_out/r2c4
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "rc=$?" >> $O/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config cfg2 --loop --keyframes 3 > $O/loop_2gpu.json 2> $O/loop_2gpu.err
echo "rc=$?" >> $O/loop_2gpu.err
tail -n 8 $O/bench_2gpu.err $O/loop_2gpu.err
