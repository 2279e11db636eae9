#!/bin/bash
# GPU call 2 of round 2: the new mapper tests (single GPU + two ranks sharing it), the drop-in test, the f32x2 microbenchmark.
set -u
O=gpurun_out/r2c2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_mapper.py tests/test_gpu_mapper_multi.py tests/test_gpu_dropin.py -m gpu -q --timeout 600 -rA > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
timeout 120 gaussian_lic_b200/_build/mb_f32x2 > $O/f32x2.txt 2>&1
tail -5 $O/pytest.log
