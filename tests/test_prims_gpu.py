"""Per-class device primitives (what the C++ facades call) against the oracle's restatement of
the same reference classes, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from sdrreceiver_b200 import binding as B

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_nco_mix_matches_oscillator_stream_order():
    import torch
    fs, f, n = 48000, 1234.0, 50000                    # crosses the table wrap at 48000
    L = B.lib()
    tab = np.zeros(2 * fs, np.float32)
    assert L.sdrb_nco_table(float(fs), f, _p(tab), fs) == fs
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, n, 2)).astype(np.float32)
    osc = np.zeros(2 * n, np.float32)
    O.lib().orc_oscillator(float(fs), f, _p(osc), n)
    want = osc.view(np.complex64)[None, :] * x.view(np.complex64)[..., 0]
    d_tab, d_x = dev(tab), dev(x)
    d_y = torch.empty_like(d_x)
    assert L.sdrb_nco_mix(d_tab.data_ptr(), fs, 0, d_x.data_ptr(), d_y.data_ptr(), 3, n, None) == 0
    got = d_y.cpu().numpy().view(np.complex64)[..., 0]
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    # continuing a stream: n0 > 0 has no start-up quirk
    assert L.sdrb_nco_mix(d_tab.data_ptr(), fs, 7, d_x.data_ptr(), d_y.data_ptr(), 3, 100, None) == 0
    got7 = d_y.cpu().numpy().view(np.complex64)[..., 0].reshape(-1)[:100]
    t = tab.view(np.complex64)
    assert np.allclose(got7, t[7:107] * x.view(np.complex64)[..., 0].reshape(-1)[:100], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("block", [12, 64, 3000])
def test_halfband11_blocks_and_carry(block):
    import torch
    L = B.lib()
    nblocks, n_ch = 4, 3
    rng = np.random.default_rng(6)
    x = rng.standard_normal((n_ch, nblocks, block, 2)).astype(np.float32)
    want = np.zeros((n_ch, nblocks * block // 2, 2), np.float32)
    for c in range(n_ch):
        O.lib().orc_halfband(_p(np.ascontiguousarray(x[c])), block, nblocks, _p(want[c]))
    hist = torch.zeros((n_ch, 11, 2), dtype=torch.float32, device="cuda")
    got = []
    for b in range(nblocks):
        d_in = dev(x[:, b])
        d_out = torch.empty((n_ch, block // 2, 2), dtype=torch.float32, device="cuda")
        assert L.sdrb_halfband11(d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, block, None) == 0
        got.append(d_out.cpu().numpy())
    got = np.concatenate(got, axis=1)
    assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    assert L.sdrb_halfband11(d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, 7, None) == -1


@pytest.mark.parametrize("taps", [11, 23, 51, 15])
@pytest.mark.parametrize("block", [8, 64, 3000])
def test_halfband_other_lengths_blocks_and_carry(taps, block):
    import torch
    L = B.lib()
    nblocks, n_ch = 4, 2
    rng = np.random.default_rng(taps + block)
    x = rng.standard_normal((n_ch, nblocks, block, 2)).astype(np.float32)
    want = np.zeros((n_ch, nblocks * block // 2, 2), np.float32)
    for c in range(n_ch):
        O.lib().orc_halfband_n(taps, _p(np.ascontiguousarray(x[c])), block, nblocks, _p(want[c]))
    hist = torch.zeros((n_ch, taps, 2), dtype=torch.float32, device="cuda")
    got = []
    for b in range(nblocks):
        d_in = dev(x[:, b])
        d_out = torch.full((n_ch, block // 2, 2), 7.0, dtype=torch.float32, device="cuda")
        assert L.sdrb_halfband(taps, d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, block, None) == 0
        got.append(d_out.cpu().numpy())
    got = np.concatenate(got, axis=1)
    if taps == 11 and block >= 12:
        assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max())   # the fused-multiply production kernel
    else:
        assert np.array_equal(got, want)                                         # same float operations, same order
    assert L.sdrb_halfband(24, d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, block, None) == -1


@pytest.mark.parametrize("ntaps,decim", [(47, 1), (49, 5), (73, 6)])
def test_fir_blocks_and_carry(ntaps, decim):
    import torch
    L = B.lib()
    n_ch, block, nblocks = 2, 600, 3
    rng = np.random.default_rng(7)
    taps = rng.standard_normal(ntaps).astype(np.float32)
    x = rng.standard_normal((n_ch, nblocks * block)).astype(np.float32)
    want = np.zeros((n_ch, nblocks * block // decim), np.float32)
    for c in range(n_ch):
        O.lib().orc_fir(ntaps, _p(taps), _p(np.ascontiguousarray(x[c])), x.shape[1], decim, _p(want[c]))
    hist = torch.zeros((n_ch, ntaps), dtype=torch.float32, device="cuda")
    d_taps = dev(taps)
    got = []
    for b in range(nblocks):
        d_in = dev(x[:, b * block:(b + 1) * block])
        d_out = torch.empty((n_ch, block // decim), dtype=torch.float32, device="cuda")
        assert L.sdrb_fir(d_taps.data_ptr(), ntaps, d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, block,
                          decim, None) == 0
        got.append(d_out.cpu().numpy())
    got = np.concatenate(got, axis=1)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_usb_demod_blocks_and_carry():
    import torch
    L = B.lib()
    n_ch, block, nblocks, fs = 2, 500, 3, 12000
    rng = np.random.default_rng(8)
    x = rng.standard_normal((n_ch, nblocks * block, 2)).astype(np.float32)
    want = np.zeros((n_ch, nblocks * block), np.float32)
    for c in range(n_ch):
        O.lib().orc_usb(125, fs, _p(np.ascontiguousarray(x[c])), nblocks * block, _p(want[c]))
    pts = np.zeros(125, np.float32)
    assert L.sdrb_hilbert_points(125, fs, _p(pts)) == 0
    d_pts = dev(pts)
    hist = torch.zeros((n_ch, 124, 2), dtype=torch.float32, device="cuda")
    got = []
    for b in range(nblocks):
        d_in = dev(x[:, b * block:(b + 1) * block])
        d_out = torch.empty((n_ch, block), dtype=torch.float32, device="cuda")
        assert L.sdrb_usb_demod(d_pts.data_ptr(), d_in.data_ptr(), d_out.data_ptr(), hist.data_ptr(), n_ch, block, None) == 0
        got.append(d_out.cpu().numpy())
    got = np.concatenate(got, axis=1)
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


@pytest.mark.skipif(not O.have_ref(), reason="kiss_fft oracle lives in oracle/_ref")
def test_spectrum_fft_against_kiss_fft():
    import torch
    L = B.lib()
    n = 8192
    rng = np.random.default_rng(9)
    x = rng.standard_normal((3, n, 2)).astype(np.float32)
    want = np.zeros_like(x)
    for k in range(3):
        O.ref_prims().ref_kiss_fft(n, _p(np.ascontiguousarray(x[k])), _p(want[k]))
    d_x = dev(x)
    d_y = torch.empty_like(d_x)
    assert L.sdrb_spectrum_fft(d_x.data_ptr(), d_y.data_ptr(), 3, n, 0, None) == 0
    got = d_y.cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-4          # SURVEY 8(d) FFT criterion
    # Hann window as in mainwindow.cpp:284-288
    hann = (0.5 * (1.0 - np.cos(2 * np.pi * np.arange(n, dtype=np.float32).astype(np.float64) / (n - 1.0)))).astype(np.float32)
    xw = (x * hann[None, :, None]).astype(np.float32)
    for k in range(3):
        O.ref_prims().ref_kiss_fft(n, _p(np.ascontiguousarray(xw[k])), _p(want[k]))
    assert L.sdrb_spectrum_fft(d_x.data_ptr(), d_y.data_ptr(), 3, n, 1, None) == 0
    got = d_y.cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-4
    assert L.sdrb_spectrum_fft(d_x.data_ptr(), d_y.data_ptr(), 3, 4096, 0, None) == -1


@pytest.mark.gpu
def test_fp32_peak_probe_is_plausible():
    """sdrb_probe_fp32_tflops: the FMA-loop peak bench.py divides by. A B200 has 148 SMs x 128 FP32 lanes;
    at 1.0-2.1 GHz that is 38-80 TFLOP/s, and a register-only FMA loop reaches most of it."""
    from sdrreceiver_b200 import binding as B
    for packed in (False, True):
        v = B.probe_fp32_tflops(packed, reps=3)
        assert 25.0 < v < 90.0, (packed, v)
