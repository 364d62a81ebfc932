#!/usr/bin/env python
"""Dynamic opcode mix, stall samples and shared-memory wavefronts per opcode from an
`ncu -i X.ncu-rep --page source --csv` export (one kernel)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
idx = {k: i for i, k in enumerate(rows[h])}
ex, st, wf = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[h + 1:]:
    if len(r) < 10 or r[0] == "Address":
        continue
    src = r[idx["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src[:10]
    n = int(r[idx["Instructions Executed"]] or 0)
    ex[op] += n
    tot += n
    st[op] += int(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    wf[op] += int(r[idx["L1 Wavefronts Shared"]] or 0)
print("total warp instructions", tot, " stall samples", sum(st.values()), " smem wavefronts", sum(wf.values()))
for op, n in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 24):
    print("%-10s %6.2f%%  stall-samples %6d  smem-wavefronts %d" % (op, 100 * n / tot, st[op], wf[op]))
