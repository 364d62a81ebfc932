#!/usr/bin/env python
"""DRAM bytes per launch and per kernel class from an `ncu -i X.ncu-rep --page raw --csv` export:
the `traffic` figure of bench.py's roofline (profiles/r01_traffic.json).
usage: ncu_traffic.py raw.csv "source note" [callbacks per call] > profiles/r02_traffic.json
Since round 2 the sub-VFO cascade, /late and USB-audio kernels are launched once per CALL (all callbacks of a process call in one grid):
their per-launch bytes are divided by the callbacks per call (third argument, default 4) so that every class is per callback."""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram(r):
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[m]]) * scale.get(units[idx[m]], 1.0)
    return tot


per = collections.defaultdict(list)
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("sdrb::", "")
    per["%s grid=%s" % (name, r[idx["launch__grid_size"]])].append(dram(r))
cls = {"k2a_v2": "sub_cascade", "k2a_v3": "sub_cascade", "k2b_v2": "usb_audio", "k1_v2": "ingest_main", "k0_dc": "dc_scan",
       "k2_late": "late_fir"}
ncb = int(sys.argv[3]) if len(sys.argv) > 3 else 4
per_call = ("k2a_v2", "k2a_v3", "k2b_v2", "k2_late")
per_class = collections.defaultdict(float)
out = {}
for k, v in per.items():
    div = ncb if k.startswith(per_call) else 1
    out[k] = {"launches_captured": len(v), "dram_bytes_per_launch": sum(v) / len(v), "callbacks_per_launch": div}
    for pre, c in cls.items():
        if k.startswith(pre):
            per_class[c] += sum(v) / len(v) / div    # per callback: one launch of each distinct kernel/grid, per-call launches divided
print(json.dumps({
    "source": sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none",
    "config": {"plan": "25E", "streams_per_gpu": 128,
               "unit": "dram__bytes_read.sum + dram__bytes_write.sum per launch; a kernel class = the launches of one callback "
                       "(128 streams x 384000 samples)"},
    "per_kernel": out, "per_class_per_callback": dict(per_class)}, indent=1))
