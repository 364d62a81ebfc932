#!/usr/bin/env python
"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): every kernel of the hot path on
short inputs, both entry points, forwarder and spectrum included. Usage:
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py quick"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from sdrreceiver_b200 import binding as B  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
rng = np.random.default_rng(1)
for name, n_streams, n_blocks in (("54W_288K", 3, 2), ("FWD_test", 2, 1)) if quick else (
        ("54W_288K", 5, 3), ("25E", 3, 2), ("54W_all", 2, 2), ("FWD_test", 2, 2)):
    plan = B.Plan(os.path.join(ROOT, "plans", name + ".ini"))
    bank = B.Bank(plan, n_streams, n_blocks)
    iq = rng.integers(100, 156, size=(n_streams, n_blocks * plan.block * 2), dtype=np.uint8)
    for rep in range(2):
        pcm, tap = bank.process_numpy(iq, n_blocks, want_tap=True)
    for k, m in enumerate(plan.mains):
        if m["forward"]:
            bank.read_forward(k, n_blocks)
    sp = B.Spectrum(n_streams)
    bank.spectrum_feed(sp, -1, 0)
    if plan.subs:
        bank.spectrum_feed(sp, 0, n_blocks - 1)
        bank.read_sub(0, n_blocks)
    bank.read_input(0, 1024)
    sp.read()
    sp.close()
    bank.close()
    print(name, "ok", int(np.abs(pcm).max()))
