#!/bin/bash
# GPU call 3 of round 2: the rewritten bench (mapper-based value leg, C++ e2e host), the reference arm, the full loop, the drop-in test.
set -u
O=gpurun_out/r2c3
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err
echo "ours rc=$?" >> $O/bench_ours.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
echo "ref rc=$?" >> $O/bench_ref.err
timeout 600 python bench.py --config cfg2 --loop --keyframes 3 > $O/loop_cfg2.json 2> $O/loop_cfg2.err
echo "loop rc=$?" >> $O/loop_cfg2.err
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q --timeout 600 > $O/pytest_dropin.log 2>&1
tail -3 $O/pytest_dropin.log
tail -5 $O/*.err
