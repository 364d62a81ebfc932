// C entry points around the reference's own DSP classes (test infrastructure).
// Built by oracle/Makefile into oracle/_ref/libref_prims.so from the unmodified
// sources under /root/reference; tests/ use it through ctypes to pin the C
// restatement (oracle/sdr_oracle.c) and the CUDA primitives class by class.
#include <cstring>
#include <vector>
#include "oscillator.h"
#include "halfbanddecimator.h"
#include "gnuradio/firfilter.h"
extern "C" {
#include "kiss_fft130/kiss_fft.h"
}

extern "C" {

// Oscillator (oscillator.cpp:4-50): the value used for stream sample i, i = 0..n-1.
void ref_oscillator(double fs, double f, float *out_iq, long n) {
    Oscillator o(fs, f);
    for (long i = 0; i < n; i++) {
        out_iq[2 * i] = o._vector.real();
        out_iq[2 * i + 1] = o._vector.imag();
        o.tick();
    }
}

// HalfBandDecimator (halfbanddecimator.cpp:43-72), nblocks calls of `block` samples.
void ref_halfband(int taps, int inlen, const float *in_iq, int block, int nblocks, float *out_iq) {
    HalfBandDecimator hb(taps, inlen);
    std::vector<cpx_typef> in(block), out(block / 2);
    for (int b = 0; b < nblocks; b++) {
        for (int i = 0; i < block; i++)
            in[i] = cpx_typef(in_iq[2 * ((long)b * block + i)], in_iq[2 * ((long)b * block + i) + 1]);
        hb.decimate(in, out);
        for (int i = 0; i < block / 2; i++) {
            out_iq[2 * ((long)b * (block / 2) + i)] = out[i].real();
            out_iq[2 * ((long)b * (block / 2) + i) + 1] = out[i].imag();
        }
    }
}

// FIR::FIRUpdateAndProcess (dsp.cpp:59-71); `every` > 1 reproduces the
// FIRUpdate / FIRUpdateAndProcess interleave of vfo::usb_decimdemod.
void ref_fir(int ntaps, const float *taps, const float *in, long n, int every, float *out) {
    FIR f(ntaps, 0);
    for (int i = 0; i < ntaps; i++) f.FIRSetPoint(i, taps[i]);
    long m = 0;
    for (long i = 0; i < n; i++) {
        if (every <= 1 || i % every == 0) out[m++] = f.FIRUpdateAndProcess(in[i]);
        else f.FIRUpdate(in[i]);
    }
}

void ref_hilbert_points(int len, int Fs, float *points) {
    FIRHilbert h(len, Fs);
    memcpy(points, h.points, sizeof(float) * len);
}

// usb = delay((len-1)/2)(re) - hilbert(im)   (vfo.cpp:136-137, 316-324)
void ref_usb(int len, int Fs, const float *in_iq, long n, float *out) {
    FIRHilbert h(len, Fs);
    DelayThing<float> d;
    d.setLength((len - 1) / 2);
    for (long i = 0; i < n; i++)
        out[i] = d.update_dont_touch(in_iq[2 * i]) - h.FIRUpdateAndProcess(in_iq[2 * i + 1]);
}

// firfilter::low_pass with WIN_HAMMING (firfilter.cpp:64-108). Returns ntaps, or
// -1 when the reference throws.
int ref_low_pass(double gain, double fs, double cutoff, double tw, float *taps, int maxn) {
    firfilter f;
    try {
        QVector<float> c = f.low_pass(gain, fs, cutoff, tw, firfilter::WIN_HAMMING, 0);
        int n = c.length();
        for (int i = 0; i < n && i < maxn; i++) taps[i] = c[i];
        return n;
    } catch (const std::out_of_range &) {
        return -1;
    }
}

// kiss_fft forward transform (kiss_fft130/kiss_fft.c), float build.
void ref_kiss_fft(int n, const float *in_iq, float *out_iq) {
    kiss_fft_cfg cfg = kiss_fft_alloc(n, 0, 0, 0);
    kiss_fft(cfg, (const kiss_fft_cpx *)in_iq, (kiss_fft_cpx *)out_iq);
    free(cfg);
}

}
