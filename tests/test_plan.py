"""Host logic: the product's plan compiler (C++, sdrb_plan_from_ini) against the oracle-side
restatement of mainwindow.cpp:27-235, its tables against the reference classes, error paths."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import PLANS, plan_path
from oracle import oracle as O, plan as OP
from sdrreceiver_b200 import binding as B


@pytest.mark.parametrize("name", PLANS)
def test_plan_matches_oracle_plan(name):
    p = B.Plan(plan_path(name)); q = OP.build_plan(plan_path(name))
    assert (p.fs, p.block, p.bufsplit, p.correct_dc, p.center) == (q["Fs"], q["block"], q["bufsplit"], q["dc"], q["center"])
    assert len(p.mains) == len(q["mains"]) and len(p.subs) == len(q["subs"])
    for a, b in zip(p.mains, q["mains"]):
        assert (a["mixer"], a["decim"], a["out_rate"]) == (b["mixer"], b["decim"], b["out_rate"])
        assert a["block_out"] == q["block"] >> b["decim"]
    off = 0
    for a, b in zip(p.subs, q["subs"]):
        for k in ("topic", "main", "decim", "late", "filterbw", "mixer", "Fs", "out_rate", "samples_out", "freq"):
            assert a[k] == b[k], (k, a, b)
        assert np.float32(a["gain"]) == np.float32(b["gain"])
        assert a["pcm_offset"] == off
        off += a["samples_out"]
    assert p.pcm_per_block == off
    assert abs(p.alg_bytes - (2 + 2 * sum(s["out_rate"] for s in q["subs"]) / q["Fs"])) < 1e-12


def test_known_plan_numbers():
    p = B.Plan(plan_path("25E"))
    assert (p.fs, p.block, p.bufsplit, p.correct_dc) == (1536000, 384000, 4, True)
    assert [(m["mixer"], m["decim"], m["out_rate"]) for m in p.mains] == [(484000.0, 2, 384000), (-496000.0, 3, 192000)]
    assert abs(p.alg_bytes - 3.140625) < 1e-12 and abs(p.alg_flops - 269.2) < 0.1
    assert p.subs[0]["n_lpf_taps"] == 29 and p.subs[18]["n_lpf_taps"] == 47      # filter_bandwidth 4000 @12k, 10000 @48k
    q = B.Plan(plan_path("54W_288K"))
    assert (q.block, q.bufsplit, q.correct_dc) == (57600, 5, False)
    assert q.subs[0]["late"] == 6 and q.subs[0]["n_dec_taps"] == 73 and q.subs[0]["samples_out"] == 9600
    r = B.Plan(plan_path("54W_all"))
    assert {s["n_dec_taps"] for s in r.subs} == {49} and {s["late"] for s in r.subs} == {5}
    c = B.Plan(plan_path("CBAND_143E"))
    assert sorted({s["n_lpf_taps"] for s in c.subs}) == [31, 155]


@pytest.mark.skipif(not os.path.exists("/root/reference/sample_ini/sdr_25E.ini"), reason="reference tree absent")
def test_reduced_plans_equal_reference_inis():
    for ref, mine in (("sdr_25E", "25E"), ("sdr_98W", "98W"), ("sdr_54W_all", "54W_all"),
                      ("sdr_54W_288K", "54W_288K"), ("CBAND_143E", "CBAND_143E")):
        a = B.Plan("/root/reference/sample_ini/%s.ini" % ref); b = B.Plan(plan_path(mine))
        assert a.mains == b.mains and a.subs == b.subs and (a.fs, a.block, a.correct_dc) == (b.fs, b.block, b.correct_dc)


def test_tables_bit_identical_to_oracle():
    p = B.Plan(plan_path("25E"))
    L = O.lib()
    for kind, idx, fs, f in ((0, 0, p.fs, p.mains[0]["mixer"]), (1, 0, p.subs[0]["Fs"], p.subs[0]["mixer"]),
                             (1, 20, p.subs[20]["Fs"], p.subs[20]["mixer"])):
        t = p.table(kind, idx)
        ref = np.zeros(2 * int(fs), np.float32)
        assert L.orc_oscillator_table(fs, f, ref.ctypes.data_as(C.c_void_p), int(fs)) == int(fs)
        assert np.array_equal(t.view(np.float32), ref)
    taps = np.zeros(512, np.float32)
    n = L.orc_low_pass(2, 48000, 10000, 2500, taps.ctypes.data_as(C.c_void_p), 512)
    assert np.array_equal(p.table(3, 18), taps[:n])
    pts = np.zeros(125, np.float32)
    L.orc_hilbert_points(125, p.subs[0]["samples_out"], pts.ctypes.data_as(C.c_void_p))
    assert np.array_equal(p.table(4, 0), pts)
    q = B.Plan(plan_path("54W_288K"))
    n = L.orc_low_pass(2, 288000, 24000, 9600, taps.ctypes.data_as(C.c_void_p), 512)
    assert n == 73 and np.array_equal(q.table(2, 0), taps[:73])


def test_host_table_entry_points():
    L = B.lib()
    t = np.zeros(2 * 1000, np.float32)
    assert L.sdrb_nco_table(1000.0, 50.0, t.ctypes.data_as(C.c_void_p), 1000) == 1000
    ref = np.zeros_like(t)
    O.lib().orc_oscillator_table(1000.0, 50.0, ref.ctypes.data_as(C.c_void_p), 1000)
    assert np.array_equal(t, ref)
    taps = np.zeros(64, np.float32)
    assert L.sdrb_low_pass(2, 48000, 10000, 2500, taps.ctypes.data_as(C.c_void_p), 64) == 47
    assert L.sdrb_low_pass(2, 48000, 30000, 2500, taps.ctypes.data_as(C.c_void_p), 64) == -1   # firdes sanity check
    pts = np.zeros(125, np.float32)
    assert L.sdrb_hilbert_points(125, 12000, pts.ctypes.data_as(C.c_void_p)) == 0 and pts[61] > 0.63


def test_plan_errors(tmp_path):
    with pytest.raises(B.SdrbError, match="cannot read"):
        B.Plan(str(tmp_path / "nope.ini"))
    bad = tmp_path / "bad.ini"
    bad.write_text("sample_rate=1000000\n[main_vfos]\nsize=0\n")
    with pytest.raises(B.SdrbError, match="not supported"):
        B.Plan(str(bad))
    far = tmp_path / "far.ini"      # a sub VFO outside every main VFO's passband is undefined in the reference
    far.write_text("sample_rate=1536000\ncenter_frequency=1545600000\n[main_vfos]\nsize=1\n1\\frequency=1545116000\n"
                   "1\\out_rate=384000\n[vfos]\nsize=1\n1\\frequency=1546500000\n1\\data_rate=600\n1\\gain=5\n1\\topic=VFO01\n")
    with pytest.raises(B.SdrbError, match="outside every main"):
        B.Plan(str(far))


def test_ini_grammar_quirks(tmp_path):
    """'#' lines are keys (not comments), ';' lines are comments, later duplicates win, spaces trimmed."""
    ini = tmp_path / "q.ini"
    ini.write_text("sample_rate = 288000\n#sample_rate=1536000\n;correct_dc_bias=1\ncenter_frequency=1546100000\n"
                   "[main_vfos]\nsize=7\nsize=1\n1\\frequency=1546100000\n1\\out_rate=288000\n"
                   "[vfos]\nsize=1\n1\\frequency=1546045422\n1\\gain=4\n1\\data_rate=10500\n1\\fiter_bandwidth=9\n1\\topic=VFO51\n")
    p = B.Plan(str(ini))
    assert p.fs == 288000 and not p.correct_dc and len(p.mains) == 1 and p.subs[0]["filterbw"] == 0
    assert OP.build_plan(str(ini))["subs"][0]["late"] == p.subs[0]["late"] == 6


def test_plan_from_desc_equals_ini():
    a = B.Plan(plan_path("CBAND_143E"))
    b = B.Plan.from_desc(a.fs, a.block, a.bufsplit, a.correct_dc, a.mains, a.subs)
    assert [s["samples_out"] for s in a.subs] == [s["samples_out"] for s in b.subs]
    assert np.array_equal(a.table(1, 3), b.table(1, 3)) and np.array_equal(a.table(3, 5), b.table(3, 5))


def test_operational_ini_keys_stay_readable(tmp_path):
    """Keys MainWindow reads for the device/GUI (mainwindow.cpp:51-96) are not hot-path inputs but the same
    parser hands them to the application."""
    src = open(plan_path("54W_288K")).read().replace("sample_rate=288000", "sample_rate=288000\ntuner_gain=496\nauto_start=1\n"
                                                     "remote_rtl=192.168.1.5:1234\nremote_rtl_gain_idx=14\ndisable_fft=1")
    p = tmp_path / "ops.ini"
    p.write_text(src)
    plan = B.Plan(str(p))
    assert plan.setting("tuner_gain") == "496" and plan.setting("auto_start") == "1" and plan.setting("disable_fft") == "1"
    assert plan.setting("remote_rtl") == "192.168.1.5:1234" and plan.setting("remote_rtl_gain_idx") == "14"
    assert plan.setting("vfos/1/topic") == plan.subs[0]["topic"] and plan.setting("main_vfos/size") == "1"
    assert plan.setting("no_such_key") is None and plan.setting("no_such_key", "x") == "x"
