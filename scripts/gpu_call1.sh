#!/bin/bash
# GPU call 1 of round 2: full GPU test suite (new pose parametrisation, un-gated extend / fused exchange), rotated golden
# vectors from the reference build, baseline bench line, launch list and ncu --set full captures of the HEAD kernels.
set -u
mkdir -p gpurun_out/r2c1
O=gpurun_out/r2c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/pytest_full.log 2>&1
echo "pytest_full rc=$?" >> $O/pytest_full.log
timeout 300 python tests/golden/make_golden.py gpurun_out/golden scene_rot_deg3 scene_rot_deg2 > $O/golden.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
# launch list (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --kernel-only --no-graph --steps 2 --warmup 1 > $O/launches.log 2>&1
# full captures of the shipped kernels (3rd steady-state instance of each)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'render_backward_kernel|render_forward_kernel|onesweep_pass_kernel|emit_keys_kernel|preprocess_forward_kernel|preprocess_backward_kernel|ssim_forward_kernel|ssim_backward_kernel' \
    -s 48 -c 16 -o $O/prof_head python bench.py --kernel-only --no-graph --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
ls -la $O
