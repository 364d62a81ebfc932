"""Spectrum display path (MainWindow::fftHandlerSlot, mainwindow.cpp:411-455 + kiss_fft): Hann window,
8192-point FFT, dB average with the 0.95/0.05 recursion, fftshift, 5-point smoothing.

Oracle: oracle/_ref/libref_prims.so -- the reference's own FFTWrapper<float>/kiss_fft with the slot's
arithmetic restated (MainWindow is a QMainWindow and cannot be linked headless). Tolerances
(SURVEY.md 8(d)): complex spectrum rel-L2 <= 1e-4; dB traces within 2e-3 dB absolute."""
import numpy as np
import pytest

from oracle import oracle as O
from sdrreceiver_b200 import binding as B

TOL_DB = 2e-3
needs_ref = pytest.mark.skipif(not O.have_ref(), reason="kiss_fft oracle lives in oracle/_ref")


def signal(rng, n, amp=4.0):
    t = np.arange(n)
    x = amp * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x += 20 * np.exp(2j * np.pi * 0.1234 * t) + 3 * np.exp(-2j * np.pi * 0.31 * t)
    return x.astype(np.complex64)


@needs_ref
def test_reference_spectrum_state_known_properties():
    """Pins the driver around the reference FFT: first update of a zero state is 0.05*10*log10(.),
    a constant input peaks at the centre bin after the fftshift, short inputs keep the stale tail."""
    sp = O.RefSpectrum()
    x = np.full(8192, 100 + 0j, np.complex64)
    sp.feed(x)
    smooth, pwr, out, stats = sp.get()
    assert np.argmax(pwr) == 4096 and abs(out[0].real - 100 * 4095.5) < 1.0     # sum of the Hann window = (N-1)/2
    assert abs(pwr[4096] - 0.5 * np.log10(1e5 * float(abs(out[0])) / 8192)) < 1e-9
    assert np.allclose(smooth, np.convolve(pwr, np.ones(5) / 5, "valid")[:8182], atol=1e-12)
    before = out.copy()
    sp.feed(np.zeros(100, np.complex64))                                         # only inr[0..99] is overwritten
    assert np.abs(sp.get()[2] - before).max() < 1e-2 * np.abs(before).max()
    sp.reset()
    sp.feed(np.zeros(100, np.complex64))
    assert np.all(sp.get()[1] == 0)                                              # max(...,1) -> 0 dB on silence
    sp.close()


@pytest.mark.gpu
@needs_ref
def test_spectrum_display_matches_reference():
    import torch
    rng = np.random.default_rng(11)
    n_disp = 3
    sp = B.Spectrum(n_disp)
    refs = [O.RefSpectrum() for _ in range(n_disp)]
    # a "Main" selection: full callbacks (only the first 8192 samples matter), 12 updates
    for it in range(12):
        x = np.stack([signal(rng, 9000, amp=1.0 + k) for k in range(n_disp)])
        d_x = torch.from_numpy(x.view(np.float32).reshape(n_disp, -1)).cuda()
        d_fft = torch.zeros((n_disp, 8192, 2), dtype=torch.float32, device="cuda")
        sp.feed_device(d_x.data_ptr(), 9000, 9000, d_fft.data_ptr())
        for k in range(n_disp):
            refs[k].feed(x[k])
    smooth, pwr, stats = sp.read()
    got_fft = d_fft.cpu().numpy().view(np.complex64)[..., 0]
    for k in range(n_disp):
        rs, rp, rout, rstats = refs[k].get()
        assert np.linalg.norm(got_fft[k] - rout) / np.linalg.norm(rout) <= 1e-4
        assert np.abs(pwr[k] - rp).max() <= TOL_DB and np.abs(smooth[k] - rs).max() <= TOL_DB
        assert np.abs(stats[k] - rstats).max() <= TOL_DB
    # selecting a sub VFO: state reset, then buffers shorter than the FFT (3000 samples at 12 kS/s):
    # the tail of the FFT input stays what it was (zeros after the reset, then stale data)
    sp.reset()
    for r in refs:
        r.reset()
    for it in range(5):
        n = 3000 if it != 2 else 8192 + 17                  # one long buffer in between leaves a stale tail behind
        x = np.stack([signal(rng, n, amp=0.5) for _ in range(n_disp)])
        sp.feed_numpy(x)
        for k in range(n_disp):
            refs[k].feed(x[k])
    smooth, pwr, stats = sp.read()
    for k in range(n_disp):
        rs, rp, _, rstats = refs[k].get()
        assert np.abs(pwr[k] - rp).max() <= TOL_DB and np.abs(smooth[k] - rs).max() <= TOL_DB
        assert np.abs(stats[k] - rstats).max() <= TOL_DB
    # one display reset alone
    sp.reset(1)
    assert np.all(sp.read()[1][1] == 0) and np.any(sp.read()[1][0] != 0)
    sp.close()
    for r in refs:
        r.close()


@pytest.mark.gpu
@needs_ref
def test_bank_spectrum_sources_match_the_reference():
    """BASELINE config 4: CBAND_143E with the spectrum path on -- 'Main' (every 4th callback, DC-corrected
    input) and a selected sub VFO (every callback), batched over the receivers of a bank, against the
    buffers the unmodified reference emits through fftData and its own FFT."""
    import os
    from conftest import plan_path
    from oracle import plan as OP
    from sdrreceiver_b200 import synth
    name, n_blocks, sub = "CBAND_143E", 5, 3
    op = OP.build_plan(plan_path(name)); plan = B.Plan(plan_path(name))
    iq = np.stack([synth.make_iq(op["Fs"], op["block"] * n_blocks, synth.carriers_for_plan(op["center"], op["subs"]), stream=s)
                   for s in range(2)])
    bank = B.Bank(plan, 2, n_blocks)
    bank.process_numpy(iq, n_blocks)
    sp_main, sp_sub = B.Spectrum(2), B.Spectrum(2)
    bank.spectrum_feed(sp_main, -1, 4)                              # the reference's first "Main" emission is callback 4
    for cb in range(n_blocks):
        bank.spectrum_feed(sp_sub, sub, cb)
    got_in = bank.read_input(4, 8192)
    got_z = bank.read_sub(sub, n_blocks)
    for s in range(2):
        e_main = O.run_ref(plan_path(name), iq[s], fft="Main")[3]
        e_sub = O.run_ref(plan_path(name), iq[s], fft=op["subs"][sub]["topic"])[3]
        assert [cb for cb, _, _ in e_main] == [4] and [cb for cb, _, _ in e_sub] == list(range(n_blocks))
        assert np.array_equal(got_in[s].view(np.uint32), e_main[0][2][:8192].view(np.uint32))
        z_ref = np.concatenate([x for _, _, x in e_sub])
        assert np.linalg.norm(got_z[s] - z_ref) / np.linalg.norm(z_ref) <= 1e-5
        r_main, r_sub = O.RefSpectrum(), O.RefSpectrum()
        r_main.feed(e_main[0][2])
        for _, _, x in e_sub:
            r_sub.feed(x)
        for sp, r in ((sp_main, r_main), (sp_sub, r_sub)):
            smooth, pwr, stats = sp.read()
            rs, rp, _, rstats = r.get()
            assert np.abs(pwr[s] - rp).max() <= TOL_DB and np.abs(smooth[s] - rs).max() <= TOL_DB
            assert np.abs(stats[s] - rstats).max() <= TOL_DB
            r.close()
    sp_main.close(); sp_sub.close(); bank.close()


@pytest.mark.gpu
def test_spectrum_argument_errors():
    import ctypes as C
    h = C.c_void_p()
    assert B.lib().sdrb_spectrum_create(0, 1, 4096, C.byref(h)) == -1
    assert B.lib().sdrb_spectrum_create(0, 0, 8192, C.byref(h)) == -1
