#!/usr/bin/env python
"""A/B of stream-group shapes for sdrb_bank_process_host (env SDRB_HOST_GROUPS) inside ONE process, two
rounds in alternating order, 25E x 128 streams x 4 callbacks per step; prints ms per step."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from sdrreceiver_b200 import binding as B, synth  # noqa: E402

plan = B.Plan(os.path.join(ROOT, "plans", "25E.ini"))
S, NB = 128, 4
row = NB * plan.block * 2
base = synth.make_iq(plan.fs, NB * plan.block, synth.carriers_for_plan(plan.center, plan.subs))
pin_in = B.PinnedBuffer(S * row)
h = pin_in.array.reshape(S, row)
for s in range(S):
    h[s] = np.roll(base, 2 * 977 * s)
pin_out = B.PinnedBuffer(S * NB * plan.pcm_per_block * 2)
shapes = sys.argv[1:] or ["1,1,1,1,1,1,1,1", "1,1,1,1", "4,4,3,2,1", "1,1", "3,3,3,2,2,1,1,1"]
res = {k: [] for k in shapes}
for rnd in range(3):
    for k in (shapes if rnd % 2 == 0 else shapes[::-1]):
        os.environ["SDRB_HOST_GROUPS"] = k
        bank = B.Bank(plan, S, NB)
        for _ in range(4):
            bank.process_host(pin_in.ptr, row, NB, pin_out.ptr, None)
        t0 = time.perf_counter()
        for _ in range(20):
            bank.process_host(pin_in.ptr, row, NB, pin_out.ptr, None)
        res[k].append((time.perf_counter() - t0) / 20 * 1e3)
        bank.close()
for k in shapes:
    print("%-22s" % k, " ".join("%.3f" % v for v in res[k]))
