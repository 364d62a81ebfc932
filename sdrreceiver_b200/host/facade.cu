// Implementation of the header-compatible C++ facades (sdrreceiver_b200/host/*.h) on top of
// the C ABI (include/sdrb200.h). Host logic only: every sample that is filtered, mixed or
// demodulated goes through an sdrb_* entry point, i.e. through a CUDA kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "../../include/sdrb200.h"
#include "dsp.h"
#include "firfilter.h"
#include "halfbanddecimator.h"
#include "oscillator.h"
#include "sdrj.h"
#include "vfo.h"
#include "zmqpublisher.h"

namespace sdrb_host {
void check(int rc, const char *what) {
    if (rc != SDRB_OK) throw Error(std::string(what) + ": " + sdrb_last_error());
}
static void cu(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw Error(std::string(what) + ": " + cudaGetErrorString(e));
}
void *dev_alloc(size_t bytes) {
    void *p = nullptr;
    cu(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc (the facades need a CUDA device; there is no CPU path)");
    return p;
}
void dev_free(void *p) { if (p) cudaFree(p); }
void to_dev(void *dst, const void *src, size_t bytes) { cu(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D"); }
void to_host(void *dst, const void *src, size_t bytes) { cu(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H"); }
void dev_zero(void *dst, size_t bytes) { cu(cudaMemset(dst, 0, bytes), "cudaMemset"); }
}  // namespace sdrb_host
using namespace sdrb_host;

// ------------------------------------------------------------------ Oscillator
Oscillator::Oscillator(double sampleRate, double Frequency) : queuePtr(0), d_table(nullptr) {
    length = (int)sampleRate;
    if (length < 1) throw Error("Oscillator: sample rate below 1 Hz");
    queue.resize((size_t)length);
    if (sdrb_nco_table(sampleRate, Frequency, reinterpret_cast<float *>(queue.data()), length) != length)
        throw Error("Oscillator: sdrb_nco_table failed");
    _vector = queue[(size_t)length - 1];                    // oscillator.cpp:26-30
}
Oscillator::~Oscillator() { dev_free(d_table); }
void Oscillator::tick() {
    queuePtr++;
    if (queuePtr == length) queuePtr = 0;
    _vector = queue[(size_t)queuePtr];
}
const float *Oscillator::deviceTable() {
    if (!d_table) {
        d_table = (float *)dev_alloc(sizeof(cpx_typef) * (size_t)length);
        to_dev(d_table, queue.data(), sizeof(cpx_typef) * (size_t)length);
    }
    return d_table;
}

// ------------------------------------------------------------------ HalfBandDecimator
HalfBandDecimator::HalfBandDecimator(int taps, int inlen) : d_in(nullptr), d_out(nullptr), d_hist(nullptr), cap(0), ntaps(taps) {
    if (taps < 3 || taps > 255 || !(taps & 1)) throw Error("HalfBandDecimator: odd filter length 3..255");
    (void)inlen;                                            // the reference only sizes its queue with it
    d_hist = (float *)dev_alloc(sizeof(float) * 2 * (size_t)taps);
    dev_zero(d_hist, sizeof(float) * 2 * (size_t)taps);
}
HalfBandDecimator::~HalfBandDecimator() { dev_free(d_in); dev_free(d_out); dev_free(d_hist); }
void HalfBandDecimator::decimate(const std::vector<cpx_typef> &in, std::vector<cpx_typef> &out) {
    const int n = (int)in.size();
    if (n > cap) {
        dev_free(d_in); dev_free(d_out);
        d_in = (float *)dev_alloc(sizeof(cpx_typef) * (size_t)n);
        d_out = (float *)dev_alloc(sizeof(cpx_typef) * (size_t)(n / 2 + 1));
        cap = n;
    }
    if ((int)out.size() < n / 2) throw Error("HalfBandDecimator::decimate: out must hold in.size()/2 samples");
    to_dev(d_in, in.data(), sizeof(cpx_typef) * (size_t)n);
    check(sdrb_halfband(ntaps, d_in, d_out, d_hist, 1, n, nullptr), "sdrb_halfband");
    to_host(out.data(), d_out, sizeof(cpx_typef) * (size_t)(n / 2));
}

// ------------------------------------------------------------------ FIR
FIR::FIR(int n, int queuesz) : NumberOfPoints(n), outsum(0), d_taps(nullptr), d_hist(nullptr), d_io(nullptr),
                               io_cap(0), taps_dirty(true) {
    (void)queuesz;
    if (n < 1 || n > 4096) throw Error("FIR: 1..4096 taps");
    points = new float[(size_t)n];
    for (int i = 0; i < n; i++) points[i] = 0;
    d_taps = (float *)dev_alloc(sizeof(float) * (size_t)n);
    d_hist = (float *)dev_alloc(sizeof(float) * (size_t)n);
    dev_zero(d_hist, sizeof(float) * (size_t)n);
    for (float &v : hb_hist) v = 0;
}
FIR::~FIR() { delete[] points; dev_free(d_taps); dev_free(d_hist); dev_free(d_io); }
void FIR::FIRSetPoint(int point, float value) {
    if (point < 0 || point >= NumberOfPoints) return;
    points[point] = value;
    taps_dirty = true;
}
void FIR::sync_taps() {
    if (taps_dirty) { to_dev(d_taps, points, sizeof(float) * (size_t)NumberOfPoints); taps_dirty = false; }
}
void FIR::process(const float *in, int n, float *out, int decim, bool include_newest) {
    if (n <= 0) return;
    sync_taps();
    if (2 * n > io_cap) { dev_free(d_io); d_io = (float *)dev_alloc(sizeof(float) * 2 * (size_t)n); io_cap = 2 * n; }
    to_dev(d_io, in, sizeof(float) * (size_t)n);
    check(sdrb_fir_ex(d_taps, NumberOfPoints, d_io, d_io + n, d_hist, 1, n, decim, include_newest ? 1 : 0, nullptr),
          "sdrb_fir_ex");
    to_host(out, d_io + n, sizeof(float) * (size_t)((n + decim - 1) / decim));
}
void FIR::FIRUpdate(float sig) { pending.push_back(sig); }
float FIR::FIRUpdateAndProcess(float sig) {
    // everything FIRUpdate()d since the last output goes to the device in one block; the output
    // wanted is the one at the newest sample
    pending.push_back(sig);
    std::vector<float> out(pending.size());
    process(pending.data(), (int)pending.size(), out.data(), 1);
    outsum = out.back();
    pending.clear();
    return outsum;
}
float FIR::FIRUpdateAndProcessHalfBandQueue(float sig) {
    // dsp.cpp:96-148. queue = [11-sample head | block so far]; the output is the dot product of
    // points[] with the 11 newest queue entries (odd taps of the half-band are zero). The host
    // only tracks WHICH samples enter; the arithmetic is one 1-output block of the FIR kernel
    // in its include-newest form, run on a scratch history.
    if (NumberOfPoints != 11) throw Error("FIR: the half-band queue path exists for 11 taps only");
    hbq.push_back(sig);
    const int have = (int)hbq.size();
    float scratch[12];                                       // [0..10] history (slot 0 unused), [11] newest
    scratch[0] = 0;
    for (int k = 0; k < 11; k++) {
        const int idx = have - 11 + k;                       // position in the block, < 0: head
        scratch[k + 1] = idx >= 0 ? hbq[(size_t)idx] : hb_hist[11 + idx];
    }
    sync_taps();
    if (io_cap < 16) { dev_free(d_io); d_io = (float *)dev_alloc(sizeof(float) * 16); io_cap = 16; }
    to_dev(d_io, scratch, sizeof(scratch));
    check(sdrb_fir_ex(d_taps, 11, d_io + 11, d_io + 12, d_io, 1, 1, 1, 1, nullptr), "sdrb_fir_ex");
    to_host(&outsum, d_io + 12, sizeof(float));
    return outsum;
}
void FIR::FIRUpdateQueue(float sig) { hbq.push_back(sig); }
void FIR::FIRQueueBackToFront() {
    // dsp.cpp:163-173: head = queue[qptr-1-N .. qptr-2]  (one slot early)
    const int have = (int)hbq.size();
    float next[11];
    for (int k = 0; k < 11; k++) {
        const int idx = have - 12 + k;
        next[k] = idx >= 0 ? hbq[(size_t)idx] : hb_hist[11 + idx >= 0 ? 11 + idx : 0];
    }
    memcpy(hb_hist, next, sizeof(next));
    hbq.clear();
}

// ------------------------------------------------------------------ FIRHilbert
FIRHilbert::FIRHilbert(int len, int Fs) : NumberOfPoints(len), outsum(0), d_taps(nullptr), d_hist(nullptr),
                                          d_io(nullptr), io_cap(0) {
    if (len < 1 || len > 4096) throw Error("FIRHilbert: 1..4096 taps");
    points = new float[(size_t)len];
    check(sdrb_hilbert_points(len, Fs, points), "sdrb_hilbert_points");
    d_taps = (float *)dev_alloc(sizeof(float) * (size_t)len);
    d_hist = (float *)dev_alloc(sizeof(float) * (size_t)len);
    to_dev(d_taps, points, sizeof(float) * (size_t)len);
    dev_zero(d_hist, sizeof(float) * (size_t)len);
}
FIRHilbert::~FIRHilbert() { delete[] points; dev_free(d_taps); dev_free(d_hist); dev_free(d_io); }
void FIRHilbert::process(const float *in, int n, float *out) {
    if (n <= 0) return;
    if (2 * n > io_cap) { dev_free(d_io); d_io = (float *)dev_alloc(sizeof(float) * 2 * (size_t)n); io_cap = 2 * n; }
    to_dev(d_io, in, sizeof(float) * (size_t)n);
    check(sdrb_fir_ex(d_taps, NumberOfPoints, d_io, d_io + n, d_hist, 1, n, 1, 1, nullptr), "sdrb_fir_ex");
    to_host(out, d_io + n, sizeof(float) * (size_t)n);
}
double FIRHilbert::FIRUpdateAndProcess(float sig) {
    float y;
    process(&sig, 1, &y);
    outsum = y;
    return outsum;
}

// ------------------------------------------------------------------ firfilter
firfilter::firfilter() : impl(nullptr) {}
firfilter::~firfilter() { delete impl; }
std::vector<float> firfilter::low_pass(double gain, double fs, double cutoff, double tw, win_type window_type, double beta) {
    (void)beta;
    if (window_type != WIN_HAMMING)
        throw std::out_of_range("firfilter::low_pass: only WIN_HAMMING (the window vfo.cpp uses) is built");
    float probe[1];
    const int n = sdrb_low_pass(gain, fs, cutoff, tw, probe, 0);
    if (n < 0) throw std::out_of_range(std::string("firdes check failed: ") + sdrb_last_error());
    std::vector<float> taps((size_t)n);
    sdrb_low_pass(gain, fs, cutoff, tw, taps.data(), n);
    return taps;
}
void firfilter::setTaps(std::vector<double> taps) {
    delete impl;
    impl = new FIR((int)taps.size(), 0);
    for (size_t i = 0; i < taps.size(); i++) impl->FIRSetPoint((int)i, (float)taps[i]);
}
double firfilter::filter(double in) {
    // firfilter.cpp:11-31 shifts the delay line, stores `in` last and sums coeffs[i]*xv[i]: the
    // newest sample is included. (Never called by the reference; float on the GPU, not double.)
    if (!impl) throw Error("firfilter::filter: setTaps() first");
    const float x = (float)in;
    float y = 0;
    impl->process(&x, 1, &y, 1, true);
    return (double)y;
}

// ------------------------------------------------------------------ ZmqPublisher
ZmqPublisher::ZmqPublisher() : connected(false), pub(nullptr), bindAddress("tcp://*:6002"), bind(false) {}
ZmqPublisher::~ZmqPublisher() { sdrb_publisher_close(pub); }
void ZmqPublisher::connect() {
    if (connected) return;
    // the reference carries on after a failed bind/connect (zmqpublisher.cpp:44-64); so do we,
    // publish() is then a no-op
    if (sdrb_publisher_open(bindAddress.c_str(), bind ? 1 : 0, &pub) != SDRB_OK) pub = nullptr;
    connected = true;
}
void ZmqPublisher::setAddress(std::string address) { bindAddress = address; }
void ZmqPublisher::setBind(bool b) { bind = b; }
void ZmqPublisher::publish(unsigned char *buf, uint32_t len, std::string topic, uint32_t sampleRate) {
    if (pub) sdrb_publisher_send(pub, topic.c_str(), sampleRate, buf, len);
}

// ------------------------------------------------------------------ vfo
ZmqPublisher vfo::bind_publisher;

vfo::vfo(void *) : mpVFOs(nullptr), Fs(0), zmqBind(false), decimateCount(0), outputRate(0), gain(0.01f),
                   mixer_freq(0), demodUSB(true), filterAudio(false), cstyle(0), filterbw(0), offsetbw(0),
                   scalecomp(1), samplesPerBuffer(0), lateDecimate(0), emitFFT(false), plan(nullptr), bank(nullptr) {}
vfo::~vfo() {
    if (bank) sdrb_bank_destroy(bank);
    if (plan) sdrb_plan_destroy(plan);
    if (mpVFOs) for (vfo *c : *mpVFOs) delete c;            // vfo.cpp:51-57
}
void vfo::setZmqAddress(std::string a) { zmqAddress = a; }
void vfo::setZmqTopic(std::string t) { zmqTopic = t; }
void vfo::setScaleComp(int s) { scalecomp = s; }
void vfo::setFs(int r) { Fs = r; }
void vfo::setDecimationCount(int c) { decimateCount = c; }
void vfo::setMixerFreq(double f) { mixer_freq = f; }
double vfo::getMixerFreq() { return mixer_freq; }
int vfo::getOutRate() { return (int)(Fs / (pow(2, decimateCount))); }
void vfo::setOffsetBandwidth(double bw) { offsetbw = (int)bw; }
void vfo::setFilterBandwidth(double bw) { filterbw = (int)bw; }
void vfo::setGain(float g) { gain = g; }
void vfo::setDemodUSB(bool u) { demodUSB = u; }
bool vfo::getDemodUSB() { return demodUSB; }
void vfo::setCompressonStyle(int st) { cstyle = st; }
void vfo::setFilter(bool f, int bw) { filterAudio = f; filterbw = bw; }
void vfo::setVFOs(std::vector<vfo *> *v) { mpVFOs = v; }
void vfo::fftVFOSlot(std::string topic) { emitFFT = (topic == zmqTopic); }

void vfo::init(int spb, bool bind, int late) {
    // vfo.cpp:60-176: sizes and publisher wiring; the DSP objects themselves live in the GPU plan
    // The BFO mix of vfo.cpp:103,307-313,346-352 (osc_bfo, active when offsetbw > 1) is not built: nothing in the
    // reference ever calls setOffsetBandwidth, so the branch is dead there. Refuse loudly instead of producing audio
    // without the mix.
    if (offsetbw > 1)
        throw Error("vfo::init: setOffsetBandwidth(>1) selects the BFO mix (vfo.cpp:307-313), which this library does not implement");
    samplesPerBuffer = spb;
    lateDecimate = late;
    int targetRate = (int)(Fs / (pow(2, decimateCount)));
    int samplesOut = (int)(spb / (pow(2, decimateCount)));
    if (demodUSB && late > 0) { targetRate /= late; samplesOut /= late; }
    outputRate = (uint32_t)targetRate;
    transmit_usb.assign((size_t)samplesOut, 0);
    transmit_iq.assign((size_t)(cstyle == 1 ? samplesOut : 2 * samplesOut), 0);          // vfo.cpp:143-150
    decimate[0].resize((size_t)spb);
    for (int a = 1; a < decimateCount + 1; a++) decimate[a].resize(decimate[a - 1].size() / 2);
    if (!vfo::bind_publisher.connected && bind) {
        vfo::bind_publisher.setAddress(zmqAddress);
        vfo::bind_publisher.setBind(bind);
        vfo::bind_publisher.connect();
    } else if (!bind) {
        connect_publisher.setBind(false);
        connect_publisher.setAddress(zmqAddress);
        connect_publisher.connect();
    }
    zmqBind = bind;
}

static void fill_sub_desc(sdrb_sub_desc &d, const std::string &topic, int main_idx, double mixer, int decim, int late,
                          int filterbw, float gain) {
    memset(&d, 0, sizeof(d));
    strncpy(d.topic, topic.c_str(), sizeof(d.topic) - 1);
    d.main_idx = main_idx; d.mixer_hz = mixer; d.decim = decim; d.late = late; d.filter_bw = filterbw; d.gain = gain;
}

static void fill_main_desc(sdrb_main_desc &d, double mixer, int decim, const std::string &topic, int scalecomp, int cstyle) {
    memset(&d, 0, sizeof(d));
    d.mixer_hz = mixer; d.decim = decim;
    strncpy(d.topic, topic.c_str(), sizeof(d.topic) - 1);
    d.compress_scale = scalecomp;
    d.compress_style = cstyle == 1 ? 1 : 2;                  // vfo.cpp:393: 1 = 4-bit arms, anything else int8 pairs
}

// A sub VFO must have been init()ed with exactly what its parent delivers per callback (mainwindow.cpp:134,223 keeps the
// two equal: buflen/2 >> decimateCount == main_out/bufsplit); the reference would index out of bounds otherwise (vfo.cpp:244).
static void check_tree_sizes(int Fs, int samplesPerBuffer, int decimateCount, const std::vector<vfo *> *subs, const char *who) {
    if (samplesPerBuffer <= 0 || Fs % samplesPerBuffer != 0)
        throw Error(std::string(who) + ": Fs must be a whole multiple of samplesPerBuffer (callbacks per second)");
    for (size_t i = 0; subs && i < subs->size(); i++)
        if ((*subs)[i]->getSamplesPerBuffer() != (samplesPerBuffer >> decimateCount))
            throw Error(std::string(who) + ": a sub VFO was init()ed with a samplesPerBuffer that is not its parent's samplesPerBuffer >> decimateCount");
}

void vfo::compile_tree() {
    check_tree_sizes(Fs, samplesPerBuffer, decimateCount, mpVFOs, "vfo");
    sdrb_plan_desc *d = new sdrb_plan_desc();
    memset(d, 0, sizeof(*d));
    d->sample_rate = Fs; d->block = samplesPerBuffer; d->bufsplit = Fs / samplesPerBuffer; d->correct_dc = 0;
    d->n_main = 1;
    fill_main_desc(d->mains[0], mixer_freq, decimateCount, zmqTopic, scalecomp, cstyle);
    d->n_sub = mpVFOs ? (int)mpVFOs->size() : 0;
    if (d->n_sub > SDRB_MAX_SUB) { delete d; throw Error("vfo: too many sub VFOs"); }
    for (int i = 0; i < d->n_sub; i++) {
        const vfo *c = (*mpVFOs)[(size_t)i];
        fill_sub_desc(d->subs[i], c->zmqTopic, 0, c->mixer_freq, c->decimateCount, c->lateDecimate, c->filterbw, c->gain);
    }
    const int rc = sdrb_plan_create(d, &plan);
    delete d;
    check(rc, "sdrb_plan_create");
    check(sdrb_bank_create(plan, 0, 1, 1, &bank), "sdrb_bank_create");
    sdrb_plan_info info;
    sdrb_plan_get_info(plan, &info);
    pcm_record.assign((size_t)std::max(info.pcm_per_block, 1), 0);
}

void vfo::process(const std::vector<cpx_typef> &samples) {
    if (demodUSB)
        throw Error("vfo::process on a leaf VFO: leaves are driven by their parent (vfo.cpp:253-266)");
    if ((int)samples.size() != samplesPerBuffer) throw Error("vfo::process: samples.size() must equal samplesPerBuffer");
    if (!bank) compile_tree();
    check(sdrb_bank_process_cf32_host(bank, reinterpret_cast<const float *>(samples.data()), samples.size(), 1,
                                      pcm_record.data(), nullptr), "sdrb_bank_process_cf32_host");
    check(sdrb_bank_read_main(bank, 0, 1, reinterpret_cast<float *>(decimate[decimateCount].data())), "sdrb_bank_read_main");
    if (!mpVFOs || mpVFOs->empty()) {                        // vfo.cpp:268-286: compress() + transmitData()
        check(sdrb_bank_read_forward(bank, 0, 1, reinterpret_cast<uint8_t *>(transmit_iq.data())), "sdrb_bank_read_forward");
        transmitData();
    }
    for (size_t i = 0; mpVFOs && i < mpVFOs->size(); i++) {
        vfo *c = (*mpVFOs)[i];
        sdrb_sub_info si;
        sdrb_plan_get_sub(plan, (int)i, &si);
        c->transmit_usb.assign(pcm_record.begin() + si.pcm_offset, pcm_record.begin() + si.pcm_offset + si.samples_out);
        c->transmitData();
        c->emit_selected(bank, (int)i);
    }
    if (emitFFT && fftData) fftData(decimate[decimateCount]);
}

// vfo.cpp:290-293: a leaf selected in the spectrum combo box emits its decimate[decimateCount]
void vfo::emit_selected(sdrb_bank *root_bank, int sub_idx) {
    if (!emitFFT || !fftData) return;
    std::vector<cpx_typef> &z = decimate[decimateCount];
    check(sdrb_bank_read_sub(root_bank, sub_idx, 1, reinterpret_cast<float *>(z.data())), "sdrb_bank_read_sub");
    fftData(z);
}

void vfo::transmitData() {                                   // vfo.cpp:426-453
    ZmqPublisher &p = zmqBind ? vfo::bind_publisher : connect_publisher;
    if (!demodUSB) {
        if (zmqTopic.length() > 0)
            p.publish(reinterpret_cast<unsigned char *>(transmit_iq.data()), (uint32_t)transmit_iq.size(), zmqTopic, outputRate);
        return;
    }
    p.publish(reinterpret_cast<unsigned char *>(transmit_usb.data()), (uint32_t)(transmit_usb.size() * sizeof(short)),
              zmqTopic, outputRate);
}

// ------------------------------------------------------------------ sdrj
sdrj::sdrj(void *) : publishEnabled(true), mpVFOs(nullptr), correctDC(false), emitFFT(false), count(0), plan(nullptr),
                     bank(nullptr) {
    floats.resize(256);
    for (int i = 0; i < 256; i++) floats[(size_t)i] = (float)(i - 127);   // sdr.cpp:43-49
}
sdrj::~sdrj() {
    if (bank) sdrb_bank_destroy(bank);
    if (plan) sdrb_plan_destroy(plan);
    if (mpVFOs) for (vfo *v : *mpVFOs) delete v;              // sdrj.cpp:19-28
}
void sdrj::setDCCorrection(bool c) { correctDC = c; }
void sdrj::setVFOs(std::vector<vfo *> *v) { mpVFOs = v; }
void sdrj::fftVFOSlot(std::string topic) { emitFFT = (topic == "Main"); count = 0; }

void sdrj::compile_tree(int block) {
    if (!mpVFOs || mpVFOs->empty()) throw Error("sdrj: setVFOs() first");
    sdrb_plan_desc *d = new sdrb_plan_desc();
    memset(d, 0, sizeof(*d));
    const vfo *m0 = (*mpVFOs)[0];
    for (const vfo *mv : *mpVFOs) {
        if (mv->Fs != m0->Fs || mv->samplesPerBuffer != block) { delete d; throw Error("sdrj: every main VFO must be init()ed with the callback size and Fs of the input"); }
        try { check_tree_sizes(mv->Fs, block, mv->decimateCount, mv->mpVFOs, "sdrj"); } catch (...) { delete d; throw; }
    }
    d->sample_rate = m0->Fs; d->block = block; d->bufsplit = m0->Fs / block; d->correct_dc = correctDC ? 1 : 0;
    d->n_main = (int)mpVFOs->size();
    if (d->n_main > SDRB_MAX_MAIN) { delete d; throw Error("sdrj: too many main VFOs"); }
    leaves.clear();
    for (int m = 0; m < d->n_main; m++) {
        const vfo *mv = (*mpVFOs)[(size_t)m];
        fill_main_desc(d->mains[m], mv->mixer_freq, mv->decimateCount, mv->zmqTopic, mv->scalecomp, mv->cstyle);
        if (!mv->mpVFOs) continue;
        for (vfo *c : *mv->mpVFOs) {
            if (d->n_sub >= SDRB_MAX_SUB) { delete d; throw Error("sdrj: too many sub VFOs"); }
            fill_sub_desc(d->subs[d->n_sub++], c->zmqTopic, m, c->mixer_freq, c->decimateCount, c->lateDecimate,
                          c->filterbw, c->gain);
            leaves.push_back(c);
        }
    }
    const int rc = sdrb_plan_create(d, &plan);
    delete d;
    check(rc, "sdrb_plan_create");
    check(sdrb_bank_create(plan, 0, 1, 1, &bank), "sdrb_bank_create");
    sdrb_plan_info info;
    sdrb_plan_get_info(plan, &info);
    pcm_record.assign((size_t)std::max(info.pcm_per_block, 1), 0);
    staging.assign(((size_t)block * 2 + 15) / 16 * 16, 127);
}

void sdrj::run(const unsigned char *bytes, uint32_t len) {
    const int block = (int)(len / 2);
    if (!bank) compile_tree(block);
    sdrb_plan_info info;
    sdrb_plan_get_info(plan, &info);
    if (block != info.block) throw Error("sdrj: callback length changed");
    if (bytes != staging.data()) memcpy(staging.data(), bytes, len);
    check(sdrb_bank_process_host(bank, staging.data(), staging.size(), 1, pcm_record.data(), nullptr),
          "sdrb_bank_process_host");
    for (size_t i = 0; i < leaves.size(); i++) {
        vfo *c = leaves[i];
        sdrb_sub_info si;
        sdrb_plan_get_sub(plan, (int)i, &si);
        c->transmit_usb.assign(pcm_record.begin() + si.pcm_offset, pcm_record.begin() + si.pcm_offset + si.samples_out);
        if (publishEnabled) c->transmitData();
        c->emit_selected(bank, (int)i);
    }
    for (size_t m = 0; m < mpVFOs->size(); m++) {             // childless main VFOs forward packed IQ (vfo.cpp:268-286)
        vfo *mv = (*mpVFOs)[m];
        if (mv->mpVFOs && !mv->mpVFOs->empty()) continue;
        check(sdrb_bank_read_forward(bank, (int)m, 1, reinterpret_cast<uint8_t *>(mv->transmit_iq.data())), "sdrb_bank_read_forward");
        if (publishEnabled) mv->transmitData();
    }
    // sdrj.cpp:296-303: every 4th buffer goes to the spectrum display when "Main" is selected
    if (count == 4 && emitFFT) {
        if (fftData) {
            samples.resize((size_t)block);                    // the DC-corrected `samples` of sdrj.cpp:271-294, bit-exact
            check(sdrb_bank_read_input(bank, 0, block, reinterpret_cast<float *>(samples.data())), "sdrb_bank_read_input");
            fftData(samples);
        }
        count = 0;
    }
    count++;
}

void sdrj::rtlsdr_callback(unsigned char *buf, uint32_t len) { run(buf, len); }

void sdrj::demodData(const float *data, int len) {
    // the floats are floats[byte] = byte - 127 (sdr.cpp:122-129): map them back to the bytes
    if (!bank) compile_tree(len / 2);
    if ((size_t)len > staging.size()) throw Error("sdrj::demodData: buffer longer than the first one");
    for (int i = 0; i < len; i++) {
        const int b = (int)data[i] + 127;
        staging[(size_t)i] = (unsigned char)(b < 0 ? 0 : b > 255 ? 255 : b);
    }
    run(staging.data(), (uint32_t)len);
}
