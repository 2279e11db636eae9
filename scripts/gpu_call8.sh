#!/bin/bash
set -u
O=gpurun_out/r2c8
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_mapper_multi.py tests/test_gpu_mapper.py -m gpu -q --timeout 600 -rA > $O/pytest.log 2>&1; grep -E "passed|failed|ranks vs" $O/pytest.log | tail -5
