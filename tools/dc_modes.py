"""How the DC walk handled each 128-sample block (GPU): fraction of blocks per mode and callback.
mode 0/1 = translation or integer solve, 2 = real float steps. Usage: python tools/dc_modes.py [seconds] [dc] [sigma]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import plan as OP
from sdrreceiver_b200 import binding as B, synth

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ini = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plans", "25E.ini")
op = OP.build_plan(ini); plan = B.Plan(ini)
car = synth.carriers_for_plan(op["center"], op["subs"])
kw = {}
if len(sys.argv) > 2: kw["dc"] = float(sys.argv[2])
if len(sys.argv) > 3: kw["sigma"] = float(sys.argv[3])
n_streams = 4
iq = np.stack([synth.make_iq(op["Fs"], op["block"] * 4, car, stream=s, **kw) for s in range(n_streams)])
bank = B.Bank(plan, n_streams, 4)
per = plan.block // 128
for sec in range(secs):
    bank.process_numpy(iq, 4)
    out = torch.empty((n_streams, 4 * per, 2), dtype=torch.float32, device="cuda")
    modes = torch.empty((n_streams, 4 * per), dtype=torch.uint8, device="cuda")
    bank.copy_dc_trace(4, out.data_ptr(), modes.data_ptr())
    torch.cuda.synchronize()
    m = modes.cpu().numpy()
    mi, mq = m & 15, m >> 4
    st = out.cpu().numpy()
    print("second %d: float-stepped blocks I %.3f Q %.3f per callback I %s ; state end %s" % (
        sec, (mi == 2).mean(), (mq == 2).mean(),
        np.round([(mi[:, k * per:(k + 1) * per] == 2).mean() for k in range(4)], 3), st[0, -1]))
bank.close()
