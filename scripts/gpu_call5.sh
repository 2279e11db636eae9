#!/bin/bash
# parity suite + stage timings after a kernel change
set -u
O=gpurun_out/r2c5
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_pin.py tests/test_gpu_fullsize.py tests/test_gpu_mapper.py tests/test_gpu_ops.py -m gpu -q --timeout 600 -x > $O/pytest.log 2>&1
echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --kernel-only > $O/bench_kernel_only.json 2> $O/bench.err
tail -3 $O/bench.err
GLIC_BWD_RP=4 timeout 600 python bench.py --steps 20 --warmup 5 --kernel-only > $O/bench_kernel_only_rp4.json 2>> $O/bench.err
