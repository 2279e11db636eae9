#!/bin/bash
set -u
O=gpurun_out/r2c9
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_mapper.py -m gpu -q --timeout 280 -k "plain_c_host" -s > $O/pytest.log 2>&1; tail -12 $O/pytest.log
