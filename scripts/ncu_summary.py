#!/usr/bin/env python
"""ncu raw-page CSV -> per-kernel summary (the columns DESIGN.md / profiles/README.md quote).

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python scripts/ncu_summary.py X_raw.csv > profiles/X_summary.csv
"""
import csv
import sys

COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("smsp__inst_executed.sum", "warp_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
        ("sm__inst_executed_pipe_fma.sum", "pipe_fma"), ("sm__inst_executed_pipe_alu.sum", "pipe_alu"),
        ("sm__inst_executed_pipe_xu.sum", "pipe_xu"), ("sm__inst_executed_pipe_lsu.sum", "pipe_lsu"),
        ("sm__inst_executed_pipe_fmaheavy.sum", "pipe_fmaheavy"), ("sm__inst_executed_pipe_fmalite.sum", "pipe_fmalite")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    pick = [(hdr.index(c), n) for c, n in COLS if c in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([n + ("[%s]" % units[i] if units[i] else "") for i, n in pick])
    for r in rows[2:]:
        out = []
        for i, n in pick:
            v = r[i]
            if n == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("unnamed>::", "")
            out.append(v)
        w.writerow(out)


if __name__ == "__main__":
    main(sys.argv[1])
