#!/usr/bin/env python
"""One line per bench log: value, ms per step, per-class ms (helper for reading gpurun_out/*_bench_*.log)."""
import glob
import json
import sys

for f in sorted(sum((glob.glob(a) for a in sys.argv[1:]), [])):
    line = [l for l in open(f).read().strip().splitlines() if l.startswith("{")]
    if not line:
        print(f, "NO JSON")
        continue
    d = json.loads(line[-1])
    pc = d.get("per_class") or (d.get("roofline") or {}).get("per_class") or {}
    print(f.split("/")[-1], "value %.1f" % d["value"], "ms %.3f" % d["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in pc.items()},
          "e2e", (d.get("e2e") or {}).get("value"), "parity", d.get("parity_ok"))
