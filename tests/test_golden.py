"""Oracle (C restatement) against golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py -> tests/golden/). Runs anywhere, no /root/reference needed."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import PLANS, level_for, plan_path
from oracle import oracle as O, plan as OP
from sdrreceiver_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_known_answers():
    ka = np.load(os.path.join(GOLD, "known_answers.npz"))
    L = O.lib()
    for fs, f in ((1536000, 484000), (384000, 110854), (192000, -73244)):
        v = np.zeros(2 * (fs + 16), np.float32)
        L.orc_oscillator(fs, f, _p(v), fs + 16)
        v = v.view(np.complex64)
        assert np.array_equal(v[:16], ka["osc_%d_%d_head" % (fs, f)])
        assert np.array_equal(v[fs - 8:fs + 16], ka["osc_%d_%d_wrap" % (fs, f)])
        assert np.array_equal(v[::4099], ka["osc_%d_%d_stride" % (fs, f)])
    y = np.zeros(64 * 4, np.float32)
    L.orc_halfband(_p(np.ascontiguousarray(ka["hb_in"])), 64, 4, _p(y))
    assert np.array_equal(y, ka["hb_out"])
    for fs in (3000, 12000):
        pts = np.zeros(125, np.float32)
        L.orc_hilbert_points(125, fs, _p(pts))
        assert np.array_equal(pts, ka["hilbert_%d" % fs])
    for key, args in (("lp_48k_10k", (2, 48000, 10000, 2500)), ("lp_dec5", (2, 60000, 6000, 3000)),
                      ("lp_dec6", (2, 288000, 24000, 9600)), ("lp_48k_3k", (2, 48000, 3000, 750))):
        t = np.zeros(512, np.float32)
        n = L.orc_low_pass(*args, _p(t), 512)
        assert n == ka[key].size and np.array_equal(t[:n], ka[key])
    u = np.zeros(400, np.float32)
    L.orc_usb(125, 12000, _p(np.ascontiguousarray(ka["usb_in"])), 400, _p(u))
    assert np.array_equal(u, ka["usb_out"])


def test_small_plan_full_vectors():
    g = np.load(os.path.join(GOLD, "plan_54W_288K_3blocks.npz"))
    op = OP.build_plan(plan_path("54W_288K"))
    iq = synth.make_iq(op["Fs"], op["block"] * 3, synth.carriers_for_plan(op["center"], op["subs"]), level=level_for(op))
    orc = O.Oracle(op, main_tap=True)
    orc.process(iq)
    for k, s in enumerate(op["subs"]):
        assert np.array_equal(orc.pcm(k), g["pcm_" + s["topic"]])
        assert np.array_equal(orc.tap(k)[::16], g["tap_" + s["topic"]])
    assert np.array_equal(orc.main_tap(0)[::64], g["main0"])
    orc.close()


@pytest.mark.parametrize("name", PLANS)
def test_plan_digests(name):
    dig = json.load(open(os.path.join(GOLD, "plan_digests.json")))[name]
    op = OP.build_plan(plan_path(name))
    iq = synth.make_iq(op["Fs"], op["block"] * 2, synth.carriers_for_plan(op["center"], op["subs"]), level=level_for(op))
    assert hashlib.sha256(iq.tobytes()).hexdigest() == dig["input_sha256"]     # generator is reproducible
    orc = O.Oracle(op)
    orc.process(iq)
    for k, s in enumerate(op["subs"]):
        assert hashlib.sha256(orc.pcm(k).tobytes()).hexdigest() == dig["pcm_sha256"][s["topic"]], s["topic"]
        assert dig["frames"][k] == [s["topic"][:5], s["out_rate"], 2 * s["samples_out"], 3]
    orc.close()


@pytest.mark.parametrize("name", ["25E", "54W_288K", "54W_all", "CBAND_143E"])
def test_bench_canaries(name):
    """tests/golden/canary_<plan>_1block.npz -- what bench.py's parity verdict compares the GPU with (first callback of a reset
    receiver, int16 of the unmodified reference): the restatement reproduces them bit for bit from the same synthetic input."""
    g = np.load(os.path.join(GOLD, "canary_%s_1block.npz" % name))
    op = OP.build_plan(plan_path(name))
    iq = synth.make_iq(op["Fs"], op["block"], synth.carriers_for_plan(op["center"], op["subs"]), level=level_for(op), stream=0)
    assert np.array_equal(np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8), g["input_sha256"])
    orc = O.Oracle(op)
    orc.process(iq)
    for k, s in enumerate(op["subs"]):
        assert np.array_equal(orc.pcm(k), g["pcm_" + s["topic"]]), s["topic"]
    orc.close()
