#!/bin/bash
# final single-GPU records: both bench arms, cfg4 loop, cfg3 single-GPU with 8 views per GPU
set -u
O=gpurun_out/r2final1
mkdir -p $O
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"
timeout 900 python bench.py --config cfg4 > $O/loop_cfg4_n1.json 2> $O/loop_cfg4_n1.err; echo "loop rc=$?"
timeout 900 python bench.py --config cfg3 --views-per-gpu 8 --steps 10 --warmup 3 --no-sort-bench > $O/bench_cfg3_n1_8views.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
tail -n 3 $O/*.err
