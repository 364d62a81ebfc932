#!/usr/bin/env python
"""PCIe ceiling of the box, for reading bench.py's e2e number: pinned H2D alone, D2H alone and both
directions at once, with the byte counts of one bench step (393 MB in, 224 MB out) in 1, 8 and 32 pieces."""
import json
import sys

import torch

H2D, D2H = 393_216_000, 224_256_000


def run(pieces, do_in, do_out, reps=10):
    dev = torch.device("cuda", 0)
    hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
    hout = torch.empty(D2H, dtype=torch.uint8).pin_memory()
    din = torch.empty(H2D, dtype=torch.uint8, device=dev)
    dout = torch.empty(D2H, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ci, co = H2D // pieces, D2H // pieces
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for k in range(pieces):
            if do_in:
                with torch.cuda.stream(s1):
                    din[k * ci:(k + 1) * ci].copy_(hin[k * ci:(k + 1) * ci], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    hout[k * co:(k + 1) * co].copy_(dout[k * co:(k + 1) * co], non_blocking=True)
        e1.record(s1)
        e2.record(s2)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    return best


if __name__ == "__main__":
    out = {}
    for pieces in (1, 8, 32):
        a, b, c = run(pieces, True, False), run(pieces, False, True), run(pieces, True, True)
        out["pieces_%d" % pieces] = {"h2d_ms": a, "h2d_gbs": H2D / a / 1e6, "d2h_ms": b, "d2h_gbs": D2H / b / 1e6,
                                     "both_ms": c, "both_h2d_gbs": H2D / c / 1e6}
    json.dump(out, sys.stdout, indent=1)
    print()
