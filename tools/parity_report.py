#!/usr/bin/env python
"""GPU-vs-oracle parity report for one plan (debug/inspection tool; the graded checks are
tests/test_parity_gpu.py). Usage: parity_report.py PLAN [n_blocks] [split] [n_streams]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import oracle as O, plan as OP  # noqa: E402
from sdrreceiver_b200 import binding as B, synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "25E"
    n_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    split = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    n_streams = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    ini = os.path.join(ROOT, "plans", name + ".ini")
    op = OP.build_plan(ini)
    plan = B.Plan(ini)
    level = 0.5 if any(s["late"] for s in op["subs"]) else 1.0
    car = synth.carriers_for_plan(op["center"], op["subs"])
    t0 = time.time()
    iq = np.stack([synth.make_iq(op["Fs"], op["block"] * n_blocks, car, stream=s, level=level) for s in range(n_streams)])
    print("synth %.1fs" % (time.time() - t0))
    bank = B.Bank(plan, n_streams, max(split, n_blocks - split, 1))
    pcm_parts, tap_parts = [], []
    row = op["block"] * 2
    b0 = 0
    for nb in [split, n_blocks - split]:
        if nb <= 0:
            continue
        p, t = bank.process_numpy(iq[:, b0 * row:(b0 + nb) * row], nb, want_tap=True)
        pcm_parts.append(p); tap_parts.append(t)
        b0 += nb
    pcm = np.concatenate(pcm_parts, axis=1)
    tap = np.concatenate(tap_parts, axis=1)
    print("launches per call:", bank.last_launches)
    worst = 0.0
    for s in range(n_streams):
        orc = O.Oracle(op)
        orc.process(iq[s])
        got_p = B.split_pcm(plan, pcm[s]); got_t = B.split_pcm(plan, tap[s])
        for k, sv in enumerate(op["subs"]):
            rp, rt = orc.pcm(k), orc.tap(k)
            gp, gt = got_p[sv["topic"]], got_t[sv["topic"]]
            rel = np.linalg.norm(gt - rt) / max(np.linalg.norm(rt), 1e-30)
            dmax = int(np.abs(gp.astype(np.int32) - rp.astype(np.int32)).max())
            worst = max(worst, rel)
            if s == 0 or rel > 1e-4 or dmax > 1:
                print("s%d %-6s decim=%d late=%d bw=%-5d n=%-6d relL2=%.3e max|dLSB|=%d peak=%d" % (
                    s, sv["topic"], sv["decim"], sv["late"], sv["filterbw"], rp.size, rel, dmax, np.abs(rp).max()))
        orc.close()
    print("worst relL2 %.3e" % worst)


if __name__ == "__main__":
    main()
