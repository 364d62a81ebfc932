// C entry points around the reference's own DSP classes (test infrastructure).
// Built by oracle/Makefile into oracle/_ref/libref_prims.so from the unmodified
// sources under /root/reference; tests/ use it through ctypes to pin the C
// restatement (oracle/sdr_oracle.c) and the CUDA primitives class by class.
#include <cstring>
#include <vector>
#include "oscillator.h"
#include "halfbanddecimator.h"
#include "gnuradio/firfilter.h"
#include <cmath>
#include "jonti/fftwrapper.h"
extern "C" {
#include "kiss_fft130/kiss_fft.h"
}

// The spectrum display state of MainWindow (mainwindow.h:54-72), without the widget. MainWindow
// itself is GUI-bound and cannot be linked headless, so fftHandlerSlot's arithmetic
// (mainwindow.cpp:411-455) is restated below as C++ with the same expression types; the FFT is
// the reference's own FFTWrapper<float> over its vendored kiss_fft.
struct RefSpectrum {
    int nFFT;
    FFTWrapper<float> *fft;
    QVector<cpx_typef> out, inr;
    QVector<float> hann_window;
    QVector<double> pwr, smooth_pwr;
    double maxval, aveval;
};

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


// MainWindow::MainWindow, spectrum part (mainwindow.cpp:243-252, 284-288)
void *ref_spectrum_new(int nFFT) {
    RefSpectrum *sp = new RefSpectrum();
    sp->nFFT = nFFT;
    sp->fft = new FFTWrapper<float>(nFFT, false);
    sp->out.resize(nFFT);
    sp->inr.resize(nFFT);
    sp->pwr.resize(nFFT);
    sp->smooth_pwr.resize(nFFT - 10);
    sp->hann_window.resize(nFFT);
    for (int i = 0; i < sp->hann_window.size(); i++) sp->hann_window[i] = 0.5 * (1.0 - cos(2 * M_PI * ((float)i) / (nFFT - 1.0)));
    sp->maxval = sp->aveval = 0;
    return sp;
}
void ref_spectrum_free(void *h) {
    RefSpectrum *sp = (RefSpectrum *)h;
    delete sp->fft;
    delete sp;
}
// on_comboVFO_currentIndexChanged (mainwindow.cpp:539-551)
void ref_spectrum_reset(void *h) {
    RefSpectrum *sp = (RefSpectrum *)h;
    for (int i = 0; i < sp->pwr.size(); i++) sp->pwr[i] = 0;
    for (int a = 0; a < sp->nFFT; a++) sp->inr[a] = 0;
}
// fftHandlerSlot (mainwindow.cpp:411-455): `lenth` complex samples arrive
void ref_spectrum_feed(void *h, const float *data_iq, int lenth) {
    RefSpectrum *sp = (RefSpectrum *)h;
    const cpx_typef *data = (const cpx_typef *)data_iq;
    double maxval = 0;
    double aveval = 0;
    for (int a = 0; a < sp->nFFT; a++) {
        if (a < lenth) sp->inr[a] = data[a] * sp->hann_window[a];
    }
    sp->fft->transform(sp->inr, sp->out);
    QVector<cpx_typef> &out = sp->out;
    QVector<double> &pwr = sp->pwr;
    const int nFFT = sp->nFFT;
    for (int i = 0; i < pwr.size(); i++) {
        int b = i + pwr.size() / 2;
        if (b >= pwr.size()) b = b - pwr.size();
        double val = 0;
        val = sqrt(out[i].imag() * out[i].imag() + out[i].real() * out[i].real());
        pwr[b] = pwr[b] * 0.95 + 0.05 * 10 * log10(fmax(100000.0 * abs((1.0 / nFFT) * val), 1));
        if (pwr[b] > maxval) maxval = pwr[b];
        aveval += pwr[b];
    }
    for (int i = 0; i < sp->smooth_pwr.size(); i++)
        sp->smooth_pwr[i] = (pwr[i + 4] + pwr[i + 3] + pwr[i + 2] + pwr[i + 1] + pwr[i]) / 5;
    aveval /= pwr.size();
    if ((maxval - aveval) < 10) maxval = aveval + 10.0;
    sp->maxval = maxval;
    sp->aveval = aveval;
}
void ref_spectrum_get(void *h, double *smooth, double *pwr, float *fft_out_iq, double *stats) {
    RefSpectrum *sp = (RefSpectrum *)h;
    if (smooth) memcpy(smooth, sp->smooth_pwr.data(), sizeof(double) * sp->smooth_pwr.size());
    if (pwr) memcpy(pwr, sp->pwr.data(), sizeof(double) * sp->pwr.size());
    if (fft_out_iq) memcpy(fft_out_iq, sp->out.data(), sizeof(cpx_typef) * sp->out.size());
    if (stats) { stats[0] = sp->maxval; stats[1] = sp->aveval; }
}

}
