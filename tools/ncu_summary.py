#!/usr/bin/env python
"""Markdown summary of an `ncu -i X.ncu-rep --page raw --csv` export: one row per distinct kernel
(first launch of each name+grid), the metrics the roofline discussion in DESIGN.md uses."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "CTAs"), ("launch__block_size", "thr"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st wait"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st no_inst"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st mio")]
units = rows[1]
print("| " + " | ".join(c[1] for c in cols) + " |")
print("|" + "---|" * len(cols))
seen = set()
for r in rows[2:]:
    key = (r[idx["Kernel Name"]].split("(")[0], r[idx["launch__grid_size"]])
    if key in seen:
        continue
    seen.add(key)
    out = []
    for name, _ in cols:
        v = r[idx[name]] if name in idx else ""
        if name == "Kernel Name":
            v = "`" + v.split("(")[0].replace("void ", "") + "`"
        elif name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            u = units[idx[name]]
            f = float(v) * {"Mbyte": 1, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1)
            v = "%.1f" % f
        elif name == "gpu__time_duration.sum":
            u = units[idx[name]]
            v = "%.1f" % (float(v) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1))
        else:
            try:
                v = "%.2f" % float(v) if "." in v else v
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
