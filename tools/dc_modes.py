#!/usr/bin/env python
"""How often does the DC walk have to step in float? (inspection tool)"""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
from sdrreceiver_b200 import binding as B, synth
plan = B.Plan(os.path.join(ROOT, "plans", "25E.ini"))
car = synth.carriers_for_plan(plan.center, plan.subs)
nb, S = 4, 4
base = synth.make_iq(plan.fs, plan.block * nb, car)
iq = np.stack([np.roll(base, 2 * 977 * s) for s in range(S)])
bank = B.Bank(plan, S, nb)
per = plan.block // 128
for call in range(8):
    bank.process_numpy(iq, nb)
    st = torch.empty((S, nb * per, 2), dtype=torch.float32, device="cuda")
    md = torch.empty((S, nb * per), dtype=torch.uint8, device="cuda")
    bank.copy_dc_trace(nb, st.data_ptr(), md.data_ptr())
    torch.cuda.synchronize()
    m = md.cpu().numpy()
    mi, mq = m & 15, m >> 4
    print("call %d: I stepped %.1f%%  Q stepped %.1f%%  state I %.6f Q %.6f" % (
        call, 100 * (mi == 2).mean(), 100 * (mq == 2).mean(), st[0, -1, 0].item(), st[0, -1, 1].item()))
