// Headless use of the drop-in C++ facades exactly the way the reference application uses its
// classes (MainWindow builds the tree with setters + init(), mainwindow.cpp:98-239; the
// dispatcher feeds callback buffers into sdrj). Writes what each leaf VFO would publish so
// tests/test_facade_gpu.py can compare it with the oracle. Links only against libsdrb200.so
// and the facade headers: no CUDA headers, no Qt.
//
//   facade_check tree  PLAN.ini IQ.u8 OUTDIR N_BLOCKS [SEL]  sdrj path (uint8 in, DC, all VFOs); SEL = spectrum
//                                                          selection ("Main" or a topic): fftData buffers -> fft.cf32/fft.txt
//   facade_check vfo   PLAN.ini IQ.u8 OUTDIR N_BLOCKS [K]  vfo::process on main VFO K (default 0; cf32 in)
//   facade_check prims OUTDIR                             per-class known answers
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "sdrb200.h"
#include "dsp.h"
#include "firfilter.h"
#include "halfbanddecimator.h"
#include "oscillator.h"
#include "sdrj.h"
#include "vfo.h"

static std::vector<unsigned char> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
template <class T>
static void append(const std::string &path, const T *p, size_t n) {
    FILE *f = fopen(path.c_str(), "ab");
    fwrite(p, sizeof(T), n, f);
    fclose(f);
}

static std::vector<vfo *> VFOmain;
static std::vector<vfo *> VFOsub[SDRB_MAX_MAIN];

// Plan numbers come from the library's ini compiler; the tree itself is built through the
// reference's own setter sequence.
static void build_tree(const char *ini, sdrb_plan_info &info) {
    sdrb_plan *plan = 0;
    if (sdrb_plan_from_ini(ini, &plan) != 0) { fprintf(stderr, "%s\n", sdrb_last_error()); exit(1); }
    sdrb_plan_get_info(plan, &info);
    for (int i = 0; i < info.n_main; i++) {
        sdrb_main_info m;
        sdrb_plan_get_main(plan, i, &m);
        vfo *pVFO = new vfo();
        pVFO->setFs(info.sample_rate);
        pVFO->setDecimationCount(m.decim);
        pVFO->setMixerFreq(m.mixer_hz);
        pVFO->setDemodUSB(false);
        pVFO->setCompressonStyle(1);
        if (m.compress_scale > 0) pVFO->setScaleComp(m.compress_scale);     // mainwindow.cpp:112-118
        if (m.topic[0]) { pVFO->setZmqAddress(""); pVFO->setZmqTopic(m.topic); }   // no socket in the test
        pVFO->init(info.block, false);
        pVFO->setVFOs(&VFOsub[i]);
        VFOmain.push_back(pVFO);
    }
    for (int i = 0; i < info.n_sub; i++) {
        sdrb_sub_info s;
        sdrb_plan_get_sub(plan, i, &s);
        vfo *pVFO = new vfo();
        pVFO->setZmqTopic(s.topic);
        pVFO->setZmqAddress("");                    // no socket in the test
        pVFO->setDecimationCount(s.decim);
        pVFO->setFilterBandwidth(s.filter_bw);
        pVFO->setGain(s.gain);
        pVFO->setMixerFreq(s.mixer_hz);
        pVFO->setFs(s.in_rate);
        pVFO->setCompressonStyle(1);
        pVFO->init(s.in_rate / info.bufsplit, true, s.late);
        VFOsub[s.main_idx].push_back(pVFO);
    }
    sdrb_plan_destroy(plan);
}

static int run_tree(int argc, char **argv) {
    if (argc < 6) return 2;
    sdrb_plan_info info;
    build_tree(argv[2], info);
    std::vector<unsigned char> iq = slurp(argv[3]);
    const std::string out = argv[4];
    const int n_blocks = atoi(argv[5]);
    sdrj *radio = new sdrj(0);
    radio->setVFOs(&VFOmain);
    radio->setDCCorrection(info.correct_dc != 0);
    radio->fftVFOSlot("none");
    radio->publishEnabled = false;
    static long callback = 0;
    if (argc > 6) {                                  // the combo box (mainwindow.cpp:228,261,539-541)
        const std::string sel = argv[6];
        auto sink = [out](const char *who) {
            return [out, who](const std::vector<cpx_typef> &v) {
                append(out + "/fft.cf32", v.data(), v.size());
                FILE *f = fopen((out + "/fft.txt").c_str(), "a");
                fprintf(f, "%ld %s %zu\n", callback, who, v.size());
                fclose(f);
            };
        };
        radio->fftData = sink("sdrj");
        radio->fftVFOSlot(sel);
        for (int m = 0; m < (int)VFOmain.size(); m++)
            for (vfo *leaf : VFOsub[m]) { leaf->fftData = sink("vfo"); leaf->fftVFOSlot(sel); }
    }
    const size_t len = (size_t)info.block * 2;
    std::vector<float> fl(len);
    for (int b = 0; b < n_blocks; b++) {
        callback = b;
        unsigned char *src = iq.data() + (size_t)b * len;
        if (b & 1) {                                 // odd callbacks through the float entry, like rtl_tcp (sdrj.cpp:155-162)
            for (size_t i = 0; i < len; i++) fl[i] = radio->floats.at(src[i]);
            radio->demodData(fl.data(), (int)len);
        } else {
            radio->rtlsdr_callback(src, (uint32_t)len);
        }
        for (int m = 0; m < (int)VFOmain.size(); m++) {
            for (vfo *leaf : VFOsub[m])
                append(out + "/" + leaf->getZmqTopic() + ".pcm", leaf->lastAudio().data(), leaf->lastAudio().size());
            if (VFOsub[m].empty() && !VFOmain[m]->getZmqTopic().empty())    // IQ forwarder (vfo::compress)
                append(out + "/" + VFOmain[m]->getZmqTopic() + ".iq8", VFOmain[m]->lastForward().data(), VFOmain[m]->lastForward().size());
        }
    }
    delete radio;                                    // deletes the whole tree, like sdrj::~sdrj
    return 0;
}

static int run_vfo(int argc, char **argv) {
    if (argc < 6) return 2;
    sdrb_plan_info info;
    build_tree(argv[2], info);
    std::vector<unsigned char> iq = slurp(argv[3]);
    const std::string out = argv[4];
    const int n_blocks = atoi(argv[5]);
    const int K = argc > 6 ? atoi(argv[6]) : 0;
    vfo *root = VFOmain[(size_t)K];
    std::vector<cpx_typef> samples((size_t)info.block);
    for (int b = 0; b < n_blocks; b++) {
        const unsigned char *src = iq.data() + (size_t)b * info.block * 2;
        for (int i = 0; i < info.block; i++)
            samples[(size_t)i] = cpx_typef((float)((int)src[2 * i] - 127), (float)((int)src[2 * i + 1] - 127));
        root->process(samples);
        const std::vector<cpx_typef> &tap = root->decimate[(int)lround(log2((double)info.sample_rate / root->getOutRate()))];
        append(out + "/main" + std::to_string(K) + ".cf32", tap.data(), tap.size());
        for (vfo *leaf : VFOsub[K])
            append(out + "/" + leaf->getZmqTopic() + ".pcm", leaf->lastAudio().data(), leaf->lastAudio().size());
        if (VFOsub[K].empty()) append(out + "/" + root->getZmqTopic() + ".iq8", root->lastForward().data(), root->lastForward().size());
    }
    bool threw = false;
    try { VFOsub[0][0]->process(samples); } catch (const sdrb_host::Error &) { threw = true; }
    printf("leaf_process_throws %d\n", threw ? 1 : 0);
    {
        // never accept-and-ignore: the BFO mix (vfo.cpp:307-313) is not implemented, init() must say so
        vfo v;
        v.setFs(48000); v.setDecimationCount(0); v.setOffsetBandwidth(1500);
        bool t2 = false;
        try { v.init(12000, true, 0); } catch (const sdrb_host::Error &) { t2 = true; }
        printf("bfo_init_throws %d\n", t2 ? 1 : 0);
    }
    {
        // a sub VFO init()ed with a callback size its parent does not deliver (out-of-bounds reads before the check)
        vfo *main = new vfo, *leaf = new vfo;
        std::vector<vfo *> subs{leaf};
        main->setFs(1536000); main->setDecimationCount(2); main->setDemodUSB(false); main->setCompressonStyle(1);
        main->setZmqAddress(""); main->init(384000, false);
        leaf->setFs(384000); leaf->setDecimationCount(3); leaf->setZmqTopic("BAD01"); leaf->init(128000, true);
        main->setVFOs(&subs);
        std::vector<cpx_typef> x(384000);
        bool t3 = false;
        try { main->process(x); } catch (const sdrb_host::Error &) { t3 = true; }
        printf("mismatched_sub_size_throws %d\n", t3 ? 1 : 0);
        subs.clear();
        main->setVFOs(nullptr);
        delete main; delete leaf;
    }
    return 0;
}

static int run_prims(int argc, char **argv) {
    if (argc < 3) return 2;
    const std::string out = argv[2];
    {   // Oscillator: stream order incl. start-up entry and table wrap
        Oscillator o(48000, 1234.0);
        std::vector<cpx_typef> v;
        for (int i = 0; i < 48010; i++) { v.push_back(o._vector); o.tick(); }
        append(out + "/osc.cf32", v.data(), v.size());
    }
    std::vector<float> x(2 * 64 * 3);
    unsigned s = 12345;
    for (float &f : x) { s = s * 1664525u + 1013904223u; f = (float)((int)(s >> 9) % 2001 - 1000) / 100.0f; }
    append(out + "/input.f32", x.data(), x.size());
    {   // HalfBandDecimator: 3 blocks of 64
        HalfBandDecimator hb(11, 48000);
        std::vector<cpx_typef> in(64), o(32);
        for (int b = 0; b < 3; b++) {
            for (int i = 0; i < 64; i++) in[(size_t)i] = cpx_typef(x[2 * (64 * b + i)], x[2 * (64 * b + i) + 1]);
            hb.decimate(in, o);
            append(out + "/hb.cf32", o.data(), o.size());
        }
        for (int taps : {23, 51, 15}) {                  // the other tables; 15 has no case in the reference: zeros
            HalfBandDecimator h2(taps, 48000);
            for (int b = 0; b < 3; b++) {
                for (int i = 0; i < 64; i++) in[(size_t)i] = cpx_typef(x[2 * (64 * b + i)], x[2 * (64 * b + i) + 1]);
                h2.decimate(in, o);
                append(out + "/hb" + std::to_string(taps) + ".cf32", o.data(), o.size());
            }
        }
        bool threw = false;
        try { HalfBandDecimator bad(24, 100); } catch (const sdrb_host::Error &) { threw = true; }
        printf("hb_even_throws %d\n", threw ? 1 : 0);
    }
    {   // FIR per-sample API in the usb_decimdemod pattern (process every 5th, update the others)
        firfilter filt;
        std::vector<float> taps = filt.low_pass(2, 60000, 6000, 3000, firfilter::WIN_HAMMING, 0);
        FIR fir((int)taps.size(), 0);
        for (size_t i = 0; i < taps.size(); i++) fir.FIRSetPoint((int)i, taps[i]);
        std::vector<float> y;
        for (int i = 0; i < 200; i++) {
            if (i % 5 == 0) y.push_back(fir.FIRUpdateAndProcess(x[(size_t)i]));
            else fir.FIRUpdate(x[(size_t)i]);
        }
        append(out + "/fir5.f32", y.data(), y.size());
        append(out + "/taps49.f32", taps.data(), taps.size());
        bool threw = false;
        try { filt.low_pass(2, 48000, 30000, 100, firfilter::WIN_HAMMING, 0); } catch (const std::out_of_range &) { threw = true; }
        printf("lowpass_throws %d\n", threw ? 1 : 0);
        // half-band queue entry points of FIR (what HalfBandDecimator::decimate calls in the reference)
        static const float hb11[11] = {0.0060431029837374152f, 0, -0.049372515458761493f, 0, 0.29332944952052842f, 0.5f,
                                       0.29332944952052842f, 0, -0.049372515458761493f, 0, 0.0060431029837374152f};
        FIR q(11, 64);
        for (int i = 0; i < 11; i++) q.FIRSetPoint(i, hb11[i]);
        std::vector<float> z;
        for (int b = 0; b < 3; b++) {
            for (int i = 0; i < 64; i++) {
                if (i % 2 == 0) z.push_back(q.FIRUpdateAndProcessHalfBandQueue(x[2 * (64 * b + i)]));
                else q.FIRUpdateQueue(x[2 * (64 * b + i)]);
            }
            q.FIRQueueBackToFront();
        }
        append(out + "/hbq.f32", z.data(), z.size());
    }
    {   // FIRHilbert + DelayThing per sample: usb = delay(re) - hilbert(im)  (vfo.cpp:316-324)
        FIRHilbert h(125, 12000);
        DelayThing<float> d;
        d.setLength(62);
        std::vector<float> u;
        for (int i = 0; i < 180; i++) u.push_back(d.update_dont_touch(x[2 * (size_t)i]) - (float)h.FIRUpdateAndProcess(x[2 * (size_t)i + 1]));
        append(out + "/usb.f32", u.data(), u.size());
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    try {
        if (!strcmp(argv[1], "tree")) return run_tree(argc, argv);
        if (!strcmp(argv[1], "vfo")) return run_vfo(argc, argv);
        if (!strcmp(argv[1], "prims")) return run_prims(argc, argv);
    } catch (const std::exception &e) {
        fprintf(stderr, "facade_check: %s\n", e.what());
        return 1;
    }
    return 2;
}
