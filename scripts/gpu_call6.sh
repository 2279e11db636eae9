#!/bin/bash
set -u
O=gpurun_out/r2c6
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity.py tests/test_gpu_model_step.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench.err
tail -3 $O/bench.err
