"""The header-compatible C++ facades (sdrreceiver_b200/host/*.h), used the way the reference
application uses its classes, against the oracle. tests/cpp/facade_check.cpp is plain C++
linked with libsdrb200.so (built by __graft_entry__.build())."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import level_for, plan_path
from oracle import oracle as O, plan as OP
from sdrreceiver_b200 import synth

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "facade_check")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run(*args):
    r = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return dict(line.split() for line in r.stdout.splitlines() if line.strip())


@pytest.mark.parametrize("name,n_blocks", [("CBAND_143E", 3), ("54W_288K", 4)])
def test_sdrj_tree_matches_oracle(name, n_blocks, tmp_path):
    op = OP.build_plan(plan_path(name))
    iq = synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]), level=level_for(op))
    iq.tofile(tmp_path / "iq.u8")
    run("tree", plan_path(name), tmp_path / "iq.u8", tmp_path, n_blocks)
    orc = O.Oracle(op)
    orc.process(iq)
    for k, s in enumerate(op["subs"]):
        got = np.fromfile(tmp_path / (s["topic"] + ".pcm"), dtype=np.int16)
        want = orc.pcm(k)
        assert got.size == want.size
        assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1, s["topic"]
    orc.close()


def test_vfo_process_on_a_main_vfo(tmp_path):
    """vfo::process(cf32) on a tree root: no DC stage, one main VFO and its children."""
    op = OP.build_plan(plan_path("25E"))
    n_blocks = 2
    iq = synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]))
    iq.tofile(tmp_path / "iq.u8")
    out = run("vfo", plan_path("25E"), tmp_path / "iq.u8", tmp_path, n_blocks)
    assert out["leaf_process_throws"] == "1"
    assert out["bfo_init_throws"] == "1" and out["mismatched_sub_size_throws"] == "1"
    sub = dict(op, dc=False, mains=[op["mains"][0]], subs=[dict(s) for s in op["subs"] if s["main"] == 0])
    orc = O.Oracle(sub, main_tap=True)
    orc.process(iq)
    tap = np.fromfile(tmp_path / "main0.cf32", dtype=np.complex64)
    ref = orc.main_tap(0)
    assert np.linalg.norm(tap - ref) / np.linalg.norm(ref) <= 1e-5
    for k, s in enumerate(sub["subs"]):
        got = np.fromfile(tmp_path / (s["topic"] + ".pcm"), dtype=np.int16)
        assert np.abs(got.astype(np.int32) - orc.pcm(k).astype(np.int32)).max() <= 1, s["topic"]
    orc.close()


def _nibbles_close(got, want):
    """vfo::compress bytes from main-VFO outputs that agree to ~1e-6: a nibble may differ by one step
    (mod 16) where a sample sits on a quantisation step, in a tiny fraction of the bytes."""
    assert got.size == want.size
    dre = ((got >> 4).astype(np.int32) - (want >> 4).astype(np.int32)) % 16
    dim = ((got & 15).astype(np.int32) - (want & 15).astype(np.int32)) % 16
    assert np.isin(dre, (0, 1, 15)).all() and np.isin(dim, (0, 1, 15)).all()
    assert np.mean(got != want) < 2e-3


def test_iq_forwarder_through_sdrj_and_vfo(tmp_path):
    """Main VFOs without sub VFOs publish vfo::compress bytes (vfo.cpp:268-286, 389-424)."""
    name, n_blocks = "FWD_test", 3
    op = OP.build_plan(plan_path(name))
    iq = synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]))
    iq.tofile(tmp_path / "iq.u8")
    (tmp_path / "tree").mkdir(); (tmp_path / "vfo").mkdir()
    run("tree", plan_path(name), tmp_path / "iq.u8", tmp_path / "tree", n_blocks)
    orc = O.Oracle(op, main_tap=True)
    orc.process(iq)
    for k in (1, 2):
        m = op["mains"][k]
        got = np.fromfile(tmp_path / "tree" / (m["topic"] + ".iq8"), dtype=np.uint8)
        _nibbles_close(got, O.compress(orc.main_tap(k), m["scalecomp"], 1))
    for k, s in enumerate(op["subs"]):
        got = np.fromfile(tmp_path / "tree" / (s["topic"] + ".pcm"), dtype=np.int16)
        assert np.abs(got.astype(np.int32) - orc.pcm(k).astype(np.int32)).max() <= 1
    orc.close()
    # vfo::process on the childless root itself (cf32 in, no DC stage, a plan with zero sub VFOs)
    run("vfo", plan_path(name), tmp_path / "iq.u8", tmp_path / "vfo", n_blocks, 1)
    solo = dict(op, dc=False, mains=[op["mains"][1]], subs=[])
    orc = O.Oracle(solo, main_tap=True)
    orc.process(iq)
    tap = np.fromfile(tmp_path / "vfo" / "main1.cf32", dtype=np.complex64)
    assert np.linalg.norm(tap - orc.main_tap(0)) / np.linalg.norm(orc.main_tap(0)) <= 1e-5
    got = np.fromfile(tmp_path / "vfo" / "IQ002.iq8", dtype=np.uint8)
    assert np.array_equal(got, O.compress(tap, 16, 1))
    _nibbles_close(got, O.compress(orc.main_tap(0), 16, 1))
    orc.close()


def _read_emits(d):
    data, at, out = np.fromfile(d / "fft.cf32", dtype=np.complex64), 0, []
    for line in open(d / "fft.txt"):
        cb, who, n = line.split()
        out.append((int(cb), who, data[at:at + int(n)]))
        at += int(n)
    return out


@pytest.mark.skipif(not O.have_ref(), reason="needs oracle/_ref (the reference's own fftData emissions)")
def test_spectrum_signals_like_the_reference(tmp_path):
    """The fftData signals (what MainWindow::fftHandlerSlot receives): 'Main' -> sdrj emits its DC-corrected
    samples on callbacks 4, 8, ... (sdrj.cpp:296-303), a sub VFO topic -> that VFO emits
    decimate[decimateCount] on every callback (vfo.cpp:290-293)."""
    name, n_blocks = "CBAND_143E", 9
    op = OP.build_plan(plan_path(name))
    iq = synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]), level=level_for(op))
    iq.tofile(tmp_path / "iq.u8")
    for sel in ("Main", op["subs"][3]["topic"]):
        d = tmp_path / sel
        d.mkdir()
        run("tree", plan_path(name), tmp_path / "iq.u8", d, n_blocks, sel)
        got = _read_emits(d)
        want = O.run_ref(plan_path(name), iq, fft=sel)[3]
        assert [(cb, who, x.size) for cb, who, x in got] == [(cb, who, x.size) for cb, who, x in want]
        for (_, _, g), (_, _, w) in zip(got, want):
            if sel == "Main":
                assert np.array_equal(g.view(np.uint32), w.view(np.uint32))        # DC-corrected input: bit-identical
            else:
                assert np.linalg.norm(g - w) / np.linalg.norm(w) <= 1e-5


def test_class_facades(tmp_path):
    out = run("prims", tmp_path)
    assert out == {"hb_even_throws": "1", "lowpass_throws": "1"}
    L = O.lib()
    x = np.fromfile(tmp_path / "input.f32", dtype=np.float32)
    # Oscillator: bit-identical, including the start-up entry and the wrap
    got = np.fromfile(tmp_path / "osc.cf32", dtype=np.float32)
    want = np.zeros(2 * 48010, np.float32)
    L.orc_oscillator(48000.0, 1234.0, _p(want), 48010)
    assert np.array_equal(got, want)
    # HalfBandDecimator::decimate, three callbacks with the off-by-one carry
    got = np.fromfile(tmp_path / "hb.cf32", dtype=np.float32)
    want = np.zeros(2 * 96, np.float32)
    L.orc_halfband(_p(x), 64, 3, _p(want))
    assert np.abs(got - want).max() <= 1e-5
    # ... and the 23/51-tap tables; a length without a case in the reference (15) filters to zeros
    for taps in (23, 51, 15):
        got = np.fromfile(tmp_path / ("hb%d.cf32" % taps), dtype=np.float32)
        want = np.zeros(2 * 96, np.float32)
        L.orc_halfband_n(taps, _p(x), 64, 3, _p(want))
        assert np.abs(got - want).max() <= 1e-5 and (taps != 15 or not got.any())
    # FIR per-sample API in the /5 pattern, taps from firfilter::low_pass
    taps = np.fromfile(tmp_path / "taps49.f32", dtype=np.float32)
    ref_taps = np.zeros(64, np.float32)
    assert L.orc_low_pass(2, 60000, 6000, 3000, _p(ref_taps), 64) == taps.size == 49
    assert np.array_equal(taps, ref_taps[:49])
    got = np.fromfile(tmp_path / "fir5.f32", dtype=np.float32)
    want = np.zeros(40, np.float32)
    L.orc_fir(49, _p(taps), _p(np.ascontiguousarray(x[:200])), 200, 5, _p(want))
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
    # FIR half-band queue entry points == the real arm of HalfBandDecimator
    got = np.fromfile(tmp_path / "hbq.f32", dtype=np.float32)
    xin = np.zeros(2 * 192, np.float32); xin[0::2] = x[0:384:2]
    want = np.zeros(2 * 96, np.float32)
    L.orc_halfband(_p(xin), 64, 3, _p(want))
    assert np.abs(got - want[0::2]).max() <= 1e-5
    # FIRHilbert + DelayThing per sample
    got = np.fromfile(tmp_path / "usb.f32", dtype=np.float32)
    want = np.zeros(180, np.float32)
    L.orc_usb(125, 12000, _p(np.ascontiguousarray(x[:360])), 180, _p(want))
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
