#!/bin/bash
# evidence of the HEAD kernels: launch list (shares) + ncu --set full of every hot kernel (3rd steady-state instance)
set -u
O=gpurun_out/r2prof
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file $O/launches.csv \
    python bench.py --kernel-only --steps 2 --warmup 3 > $O/launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'render_backward_kernel|render_forward_kernel|onesweep_pass_kernel|emit_keys_kernel|preprocess_forward_kernel|preprocess_backward_kernel|ssim_forward_kernel|ssim_backward_kernel|adam_compact_kernel|depth_scan_kernel|tile_ranges_kernel' \
    -s 60 -c 22 -o $O/prof_head python bench.py --kernel-only --steps 2 --warmup 3 > $O/ncu_full.log 2>&1
ncu -i $O/prof_head.ncu-rep --page raw --csv > $O/prof_head_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/prof_head_raw.csv > $O/prof_head_summary.csv
cat $O/prof_head_summary.csv | cut -c1-200
