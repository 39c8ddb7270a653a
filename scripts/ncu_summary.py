#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a few lines: python scripts/ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math_thr"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio_thr"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_not_sel"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg_thr"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:60]
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = "%.4g" % float(v)
                except ValueError:
                    pass
                parts.append(f"{short}={v}{units[i] if short in ('dram_rd','dram_wr') else ''}")
        print(f"{f}: {name}\n    " + " ".join(parts))
