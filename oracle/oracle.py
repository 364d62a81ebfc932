"""ctypes front end of the oracle (test infrastructure).

* `Oracle`      -- oracle/sdr_oracle.c, the plain-C restatement (one object = one stream)
* `run_ref`     -- oracle/_ref/sdr_ref_{i16,f32}, the unmodified reference hot path
* `ref_prims()` -- oracle/_ref/libref_prims.so, the reference's own DSP classes

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_lib = None
_prims = None


def build(force=False):
    """make -C oracle (C restatement always; _ref only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "libsdr_oracle.so")) \
            or (os.path.exists("/root/reference/vfo.cpp") and not have_ref()):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)


def have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f))
               for f in ("sdr_ref_i16", "sdr_ref_f32", "libref_prims.so"))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "libsdr_oracle.so"))
        vp, i, d, f, l = C.c_void_p, C.c_int, C.c_double, C.c_float, C.c_long
        L.orc_create.restype = vp; L.orc_create.argtypes = [i, i, i]
        L.orc_add_main.restype = i; L.orc_add_main.argtypes = [vp, d, i, i]
        L.orc_add_sub.restype = i; L.orc_add_sub.argtypes = [vp, i, i, d, i, i, i, i, f]
        L.orc_process.restype = None; L.orc_process.argtypes = [vp, vp, l]
        L.orc_sub_count.restype = l; L.orc_sub_count.argtypes = [vp, i]
        L.orc_sub_pcm.restype = vp; L.orc_sub_pcm.argtypes = [vp, i]
        L.orc_sub_tap.restype = vp; L.orc_sub_tap.argtypes = [vp, i]
        L.orc_main_count.restype = l; L.orc_main_count.argtypes = [vp, i]
        L.orc_main_tap.restype = vp; L.orc_main_tap.argtypes = [vp, i]
        L.orc_clear_outputs.restype = None; L.orc_clear_outputs.argtypes = [vp]
        L.orc_destroy.restype = None; L.orc_destroy.argtypes = [vp]
        L.orc_oscillator.restype = None; L.orc_oscillator.argtypes = [d, d, vp, l]
        L.orc_oscillator_table.restype = i; L.orc_oscillator_table.argtypes = [d, d, vp, l]
        L.orc_halfband.restype = None; L.orc_halfband.argtypes = [vp, i, i, vp]
        L.orc_halfband_n.restype = None; L.orc_halfband_n.argtypes = [i, vp, i, i, vp]
        L.orc_fir.restype = None; L.orc_fir.argtypes = [i, vp, vp, l, i, vp]
        L.orc_hilbert_points.restype = None; L.orc_hilbert_points.argtypes = [i, i, vp]
        L.orc_usb.restype = None; L.orc_usb.argtypes = [i, i, vp, l, vp]
        L.orc_low_pass.restype = i; L.orc_low_pass.argtypes = [d, d, d, d, vp, i]
        L.orc_dc_trace.restype = None; L.orc_dc_trace.argtypes = [vp, l, i, vp]
        L.orc_compress.restype = None; L.orc_compress.argtypes = [vp, l, i, i, vp]
        L.orc_input_samples.restype = None; L.orc_input_samples.argtypes = [vp, l, i, vp]
        _lib = L
    return _lib


def ref_prims():
    global _prims
    if _prims is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_prims.so"))
        vp, i, d, l = C.c_void_p, C.c_int, C.c_double, C.c_long
        L.ref_oscillator.restype = None; L.ref_oscillator.argtypes = [d, d, vp, l]
        L.ref_halfband.restype = None; L.ref_halfband.argtypes = [i, i, vp, i, i, vp]
        L.ref_fir.restype = None; L.ref_fir.argtypes = [i, vp, vp, l, i, vp]
        L.ref_hilbert_points.restype = None; L.ref_hilbert_points.argtypes = [i, i, vp]
        L.ref_usb.restype = None; L.ref_usb.argtypes = [i, i, vp, l, vp]
        L.ref_low_pass.restype = i; L.ref_low_pass.argtypes = [d, d, d, d, vp, i]
        L.ref_kiss_fft.restype = None; L.ref_kiss_fft.argtypes = [i, vp, vp]
        L.ref_spectrum_new.restype = vp; L.ref_spectrum_new.argtypes = [i]
        L.ref_spectrum_free.restype = None; L.ref_spectrum_free.argtypes = [vp]
        L.ref_spectrum_reset.restype = None; L.ref_spectrum_reset.argtypes = [vp]
        L.ref_spectrum_feed.restype = None; L.ref_spectrum_feed.argtypes = [vp, vp, i]
        L.ref_spectrum_get.restype = None; L.ref_spectrum_get.argtypes = [vp, vp, vp, vp, vp]
        _prims = L
    return _prims


class RefSpectrum:
    """MainWindow's spectrum state driven headless (oracle/ref_prims.cpp): the reference's FFTWrapper and
    kiss_fft, fftHandlerSlot's arithmetic restated (mainwindow.cpp:411-455)."""

    def __init__(self, nfft=8192):
        self.nfft = nfft
        self.h = ref_prims().ref_spectrum_new(nfft)

    def reset(self):
        ref_prims().ref_spectrum_reset(self.h)

    def feed(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        ref_prims().ref_spectrum_feed(self.h, _p(x.view(np.float32)), x.size)

    def get(self):
        smooth, pwr = np.zeros(self.nfft - 10), np.zeros(self.nfft)
        out, stats = np.zeros(2 * self.nfft, np.float32), np.zeros(2)
        ref_prims().ref_spectrum_get(self.h, _p(smooth), _p(pwr), _p(out), _p(stats))
        return smooth, pwr, out.view(np.complex64), stats

    def close(self):
        if self.h:
            ref_prims().ref_spectrum_free(self.h)
            self.h = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One stream of one plan through the C restatement."""

    def __init__(self, plan, main_tap=False):
        L = lib()
        self.plan = plan
        self.h = L.orc_create(plan["Fs"], plan["block"], int(plan["dc"]))
        for m in plan["mains"]:
            assert L.orc_add_main(self.h, m["mixer"], m["decim"], int(main_tap)) >= 0
        for s in plan["subs"]:
            assert L.orc_add_sub(self.h, s["main"], s["Fs"], s["mixer"], s["decim"],
                                 s["samples_per_buffer"], s["late"], s["filterbw"], s["gain"]) >= 0

    def process(self, iq_u8):
        iq = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        nblocks = iq.size // (2 * self.plan["block"])
        lib().orc_process(self.h, _p(iq), nblocks)
        return nblocks

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).copy()

    def pcm(self, s):
        L = lib()
        return self._arr(L.orc_sub_pcm(self.h, s), L.orc_sub_count(self.h, s), np.int16)

    def tap(self, s):
        L = lib()
        return self._arr(L.orc_sub_tap(self.h, s), L.orc_sub_count(self.h, s), np.float32)

    def main_tap(self, m):
        L = lib()
        n = L.orc_main_count(self.h, m)
        return self._arr(L.orc_main_tap(self.h, m), 2 * n, np.float32).view(np.complex64)

    def clear(self):
        lib().orc_clear_outputs(self.h)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compress(cf, scalecomp=1, cstyle=1):
    """vfo::compress (vfo.cpp:389-424) of complex64 samples: uint8 payload."""
    x = np.ascontiguousarray(cf, dtype=np.complex64)
    out = np.zeros(x.size * (1 if cstyle == 1 else 2), dtype=np.uint8)
    lib().orc_compress(_p(x.view(np.float32)), x.size, int(scalecomp), int(cstyle), _p(out))
    return out


def input_samples(iq_u8, correct_dc):
    """The `samples` vector of sdrj::demodData for the whole stream (sdrj.cpp:271-294): complex64."""
    iq = np.ascontiguousarray(iq_u8, dtype=np.uint8)
    n = iq.size // 2
    out = np.zeros(2 * n, dtype=np.float32)
    lib().orc_input_samples(_p(iq), n, int(bool(correct_dc)), _p(out))
    return out.view(np.complex64)


def dc_trace(iq_u8, every=32):
    """avept (sdrj.cpp:280) entering every `every`-th sample, complex64."""
    iq = np.ascontiguousarray(iq_u8, dtype=np.uint8)
    n = iq.size // 2
    out = np.zeros(2 * (n // every), dtype=np.float32)
    lib().orc_dc_trace(_p(iq), n, every, _p(out))
    return out.view(np.complex64)


def run_ref(ini_path, iq_u8, float_tap=False, main_tap=False, blocks=None, fft=None, via=None):
    """Run the unmodified reference on `iq_u8`. Returns (outputs, frames, mains):
    outputs[topic] = int16 (or float32) array, frames = list of (topic_bytes, rate,
    payload_bytes, parts), mains[k] = complex64 decimate[decimateCount] of main k.
    fft="Main" or a sub VFO topic: the combo-box selection; a 4th value is returned, the list of
    (callback, "sdrj"|"vfo", complex64 buffer) the reference emitted through its fftData signals.
    via="callback[:N]": the bytes enter through sdr::rtlsdr_callback, N callbacks queued per dispatcher run
    (jonti/sdr.cpp:100-184); via="rtltcp:S1,S2,..": through sdrj::start_tcp_rtl / sdrj::readyRead with arrivals of
    S1, S2, .. bytes (sdrj.cpp:31-74,125-166). With `via` a 4th value is returned: {"tcp_tx": bytes the client wrote,
    "ingest": the harness's one-line report as a dict}."""
    exe = os.path.join(REF_DIR, "sdr_ref_f32" if float_tap else "sdr_ref_i16")
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "iq.u8")
        np.ascontiguousarray(iq_u8, dtype=np.uint8).tofile(inp)
        cmd = [exe, "--ini", ini_path, "--in", inp, "--out", d]
        if main_tap:
            cmd.append("--main-tap")
        if blocks is not None:
            cmd += ["--blocks", str(blocks)]
        if fft is not None:
            cmd += ["--fft", fft]
        if via is not None:
            cmd += ["--via", via]
        subprocess.run(cmd, check=True)
        outs, frames, mains = {}, [], {}
        for fn in sorted(os.listdir(d)):
            if fn.endswith(".pcm"):
                outs[fn[:-4]] = np.fromfile(os.path.join(d, fn), dtype=np.float32 if float_tap else np.int16)
            elif fn.startswith("main") and fn.endswith(".cf32"):
                mains[int(fn[4:-5])] = np.fromfile(os.path.join(d, fn), dtype=np.complex64)
        with open(os.path.join(d, "frames.txt")) as f:
            for line in f:
                t, rate, nb, parts = line.split()
                frames.append((bytes.fromhex(t), int(rate), int(nb), int(parts)))
        if via is not None:
            info = {"tcp_tx": b"", "ingest": {}}
            if os.path.exists(os.path.join(d, "tcp_tx.bin")):
                with open(os.path.join(d, "tcp_tx.bin"), "rb") as f:
                    info["tcp_tx"] = f.read()
            with open(os.path.join(d, "ingest.txt")) as f:
                w = f.read().split()
                info["ingest"] = {w[i]: int(w[i + 1]) for i in range(0, len(w) - 1, 2)}
            return outs, frames, mains, info
        if fft is not None:
            emits, data, at = [], np.fromfile(os.path.join(d, "fft.cf32"), dtype=np.complex64), 0
            with open(os.path.join(d, "fft.txt")) as f:
                for line in f:
                    cb, who, n = line.split()
                    emits.append((int(cb), who, data[at:at + int(n)].copy()))
                    at += int(n)
            return outs, frames, mains, emits
    return outs, frames, mains


def time_ref(ini_path, iq_path, blocks):
    """Seconds the reference spends in byte->float + demodData for `blocks` callbacks."""
    exe = os.path.join(REF_DIR, "sdr_ref_i16")
    out = subprocess.run([exe, "--ini", ini_path, "--in", iq_path, "--time", "--blocks", str(blocks)],
                         check=True, capture_output=True, text=True).stdout.split()
    return int(out[0]), float(out[1])
