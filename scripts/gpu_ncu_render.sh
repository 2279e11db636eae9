#!/bin/bash
set -u
O=gpurun_out/r2ncu
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_backward_kernel|render_forward_kernel' -s 6 -c 2 -o $O/render_$1 python bench.py --kernel-only --steps 2 --warmup 3 > $O/ncu_$1.log 2>&1
ncu -i $O/render_$1.ncu-rep --page raw --csv > $O/render_$1_raw.csv 2>/dev/null
python scripts/ncu_summary.py $O/render_$1_raw.csv
