#!/bin/bash
set -u
O=gpurun_out/r2c7
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_model_step.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -15 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-sort-bench > $O/bench_ours.json 2> $O/bench.err
tail -3 $O/bench.err
