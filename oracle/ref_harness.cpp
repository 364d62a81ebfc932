// Headless driver for the UNMODIFIED SDRReceiver hot path (test infrastructure).
//
// oracle/Makefile compiles this file together with the reference's own
// translation units, taken where they lie under /root/reference, against the
// Qt/zmq/rtl-sdr stand-ins in oracle/shim/. The result (oracle/_ref/sdr_ref_*)
// is the ground truth every parity test and the CPU baseline are pinned to.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may execute it.
//
// What is restated here (because MainWindow is GUI-bound and cannot be linked):
//   * the QSettings ini grammar the sample plans use          (mainwindow.cpp:27)
//   * callback size and VFO-tree construction                 (mainwindow.cpp:67-239)
//   * the rtl_tcp / librtlsdr byte -> float step via sdr::floats (jonti/sdr.cpp:122-129,
//     sdrj.cpp:155-162)
// Everything downstream of sdrj::demodData() is the reference's own object code.
//
// usage: sdr_ref --ini plan.ini --in iq.u8 --out DIR [--blocks N] [--main-tap] [--time]
//   DIR/<topic>.pcm   concatenated ZMQ payloads of that topic (int16, or float32 in
//                     the REF_FLOAT_TAP build where vfo.cpp is compiled with
//                     `#define short float`)
//   DIR/frames.txt    one line per ZMQ message: topic-bytes(hex) rate payload-bytes parts
//   DIR/main<k>.cf32  (--main-tap) decimate[decimateCount] of main VFO k, every block
//   --time            no files; prints "samples seconds" for the demodData loop only
//   --skip N          (with --time) first N callbacks run untimed (warm-up)
//   --fft SEL         what MainWindow's VFO combo box does (mainwindow.cpp:539-541): SEL = "Main" or
//                     a sub VFO topic is sent to sdrj::fftVFOSlot and every sub VFO's fftVFOSlot
//                     (mainwindow.cpp:228,261); the emitted buffers go to DIR/fft.cf32 / fft.txt
#include <chrono>
#include <cstdio>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include <complex>
#include <cstring>
#include <cmath>
#include "qshim_core.h"
// The two ingest entry points of the reference are private members (sdr::rtlsdr_callback / sdr::demod_dispatcher,
// jonti/sdr.h:64-79; sdrj::readyRead / sdrj::sendCommand, a private slot and a private member of sdrj.h). The reference's
// translation units are compiled untouched; only THIS file sees their headers with the access specifiers opened, so
// that the harness can call what librtlsdr's thread, the dispatcher thread and Qt's event loop call in the application.
#define private public
#define protected public
#include "sdrj.h"
#undef private
#undef protected

// ---- moc stand-ins (signals are plain functions once Q_OBJECT is empty) ----
// With --fft SEL the two fftData signals write what they carry (the spectrum display's input,
// mainwindow.cpp:227,259) to DIR/fft.cf32 and one line "callback source length" to DIR/fft.txt.
static FILE *g_fft_out = 0, *g_fft_log = 0;
static long g_callback = 0;
static void fft_emit(const char *who, const std::vector<cpx_typef> &v) {
    if (!g_fft_out) return;
    fwrite(v.data(), sizeof(cpx_typef), v.size(), g_fft_out);
    fprintf(g_fft_log, "%ld %s %zu\n", g_callback, who, v.size());
}
void vfo::fftData(const std::vector<cpx_typef> &v) { fft_emit("vfo", v); }
void sdrj::fftData(const std::vector<cpx_typef> &v) { fft_emit("sdrj", v); }
// MainWindow connects sdr::audio_signal_out to sdrj::demodData (mainwindow.cpp:260)
static sdrj *g_audio_sink = 0;
void sdr::audio_signal_out(const float *d, int n) { if (g_audio_sink) g_audio_sink->demodData(d, n); }

// ---- librtlsdr: no device ----
extern "C" {
static bool g_have_dongle = false;
int rtlsdr_open(rtlsdr_dev_t **, uint32_t) { return g_have_dongle ? 0 : -1; }
int rtlsdr_close(rtlsdr_dev_t *) { return 0; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *) { return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *, uint32_t) { return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t *, uint32_t) { return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_agc_mode(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_set_bias_tee(rtlsdr_dev_t *, int) { return 0; }
int rtlsdr_read_async(rtlsdr_dev_t *, rtlsdr_read_async_cb_t, void *, uint32_t, uint32_t) { return 0; }
int rtlsdr_cancel_async(rtlsdr_dev_t *) { return 0; }
uint32_t rtlsdr_get_device_count(void) { return 0; }
const char *rtlsdr_get_device_name(uint32_t) { return ""; }
int rtlsdr_get_device_usb_strings(uint32_t, char *, char *, char *) { return -1; }
int rtlsdr_get_index_by_serial(const char *) { return -1; }
}

// ---- libzmq: in-memory capture of what ZmqPublisher::publish() sends ----
struct ZmqMessage { std::vector<std::string> parts; };
static std::vector<ZmqMessage> g_messages;
static ZmqMessage g_partial;
static bool g_capture = true;
extern "C" {
void *zmq_ctx_new(void) { return (void *)1; }
void *zmq_socket(void *, int) { return (void *)1; }
int zmq_setsockopt(void *, int, const void *, size_t) { return 0; }
int zmq_bind(void *, const char *) { return 0; }
int zmq_connect(void *, const char *) { return 0; }
int zmq_errno(void) { return 0; }
int zmq_send(void *, const void *buf, size_t len, int flags) {
    if (!g_capture) return (int)len;
    g_partial.parts.emplace_back((const char *)buf, len);
    if (!(flags & ZMQ_SNDMORE)) { g_messages.push_back(g_partial); g_partial.parts.clear(); }
    return (int)len;
}
}

// ---- QSettings(IniFormat) subset: flat "group/key" -> string map ----
struct Ini {
    std::map<std::string, std::string> kv;
    static std::string trim(const std::string &s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    bool load(const char *path) {
        std::ifstream f(path);
        if (!f) return false;
        std::string line, group;
        while (std::getline(f, line)) {
            line = trim(line);
            if (line.empty() || line[0] == ';') continue;   // '#' is NOT a comment for QSettings
            if (line[0] == '[') {
                size_t e = line.find(']');
                group = trim(line.substr(1, e == std::string::npos ? std::string::npos : e - 1));
                if (group == "General") group.clear();
                continue;
            }
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string k = trim(line.substr(0, eq)), v = trim(line.substr(eq + 1));
            for (char &c : k) if (c == '\\') c = '/';
            if (v.size() >= 2 && v.front() == '"' && v.back() == '"') v = v.substr(1, v.size() - 2);
            kv[group.empty() ? k : group + "/" + k] = v;
        }
        return true;
    }
    QString value(const std::string &k) const {
        auto it = kv.find(k);
        return it == kv.end() ? QString("") : QString(it->second);
    }
};

int main(int argc, char **argv) {
    const char *ini_path = 0, *in_path = 0, *out_dir = 0;
    long max_blocks = -1, skip_blocks = 0;
    bool main_tap = false, timing = false;
    const char *fft_sel = 0;
    // --via callback[:N]   bytes enter through sdr::rtlsdr_callback (librtlsdr's thread, jonti/sdr.cpp:100-145), N callbacks
    //                      (default 1) are queued before the dispatcher (sdr::demod_dispatcher, :147-184) runs until it
    //                      would sleep; more than 20 queued = the reference drops the rest ("Dropped RTL buffer!!")
    // --via rtltcp:S1,S2.. bytes enter through the rtl_tcp client (sdrj::start_tcp_rtl / sdrj::readyRead, sdrj.cpp:31-74,
    //                      125-166): the 12-byte dongle header arrives first, then the stream in pieces of S1, S2, .. bytes
    //                      (cyclic), one readyRead per arrival; what the client wrote goes to DIR/tcp_tx.bin
    std::string via;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--ini" && i + 1 < argc) ini_path = argv[++i];
        else if (a == "--in" && i + 1 < argc) in_path = argv[++i];
        else if (a == "--out" && i + 1 < argc) out_dir = argv[++i];
        else if (a == "--blocks" && i + 1 < argc) max_blocks = atol(argv[++i]);
        else if (a == "--main-tap") main_tap = true;
        else if (a == "--time") timing = true;
        else if (a == "--skip" && i + 1 < argc) skip_blocks = atol(argv[++i]);
        else if (a == "--fft" && i + 1 < argc) fft_sel = argv[++i];
        else if (a == "--via" && i + 1 < argc) via = argv[++i];
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (!ini_path || !in_path || (!out_dir && !timing)) {
        fprintf(stderr, "usage: %s --ini F --in iq.u8 --out DIR [--blocks N] [--main-tap] [--time]\n", argv[0]);
        return 2;
    }
    Ini settings;
    if (!settings.load(ini_path)) { fprintf(stderr, "cannot read %s\n", ini_path); return 1; }

    // ---- plan construction, restating mainwindow.cpp:29-239 ----
    int Fs = settings.value("sample_rate").toInt();
    if (Fs != 288000 && Fs != 1536000 && Fs != 1920000) { fprintf(stderr, "unsupported sample_rate %d\n", Fs); return 1; }
    int center_frequency = settings.value("center_frequency").toInt();
    int mix_offset = settings.value("mix_offset").toInt();
    int bufsplit = 4, buflen;
    if (double((int((2 * Fs) / 4)) % 512) > 0) { buflen = int((2 * Fs) / 5); bufsplit = 5; }
    else buflen = int((2 * Fs) / 4);
    QString zmq_address = settings.value("zmq_address");
    bool dc = settings.value("correct_dc_bias") == "1";

    static QVector<vfo *> VFOmain;
    static QVector<vfo *> VFOsub[3];
    int msize = settings.value("main_vfos/size").toInt();
    for (int i = 0; i < msize; ++i) {
        std::string p = "main_vfos/" + std::to_string(i + 1) + "/";
        vfo *pVFO = new vfo();
        int vfo_freq = settings.value(p + "frequency").toInt();
        int vfo_out_rate = settings.value(p + "out_rate").toInt();
        QString output_connect = settings.value(p + "zmq_address");
        QString out_topic = settings.value(p + "zmq_topic");
        int compscale = settings.value(p + "compress_scale").toInt();
        if (compscale > 0) pVFO->setScaleComp(compscale);
        if (output_connect != "" && out_topic != "") {
            pVFO->setZmqAddress(output_connect);
            pVFO->setZmqTopic(out_topic);
        }
        pVFO->setFs(Fs);
        pVFO->setDecimationCount(Fs / vfo_out_rate == 1 ? 0 : int(log2(Fs / vfo_out_rate)));
        pVFO->setMixerFreq(center_frequency - vfo_freq);
        pVFO->setDemodUSB(false);
        pVFO->setCompressonStyle(1);
        pVFO->init(buflen / 2, false);
        pVFO->setVFOs(&VFOsub[i]);
        VFOmain.push_back(pVFO);
    }
    int size = settings.value("vfos/size").toInt();
    std::vector<std::string> topics;
    for (int i = 0; i < size; ++i) {
        std::string p = "vfos/" + std::to_string(i + 1) + "/";
        vfo *pVFO = new vfo();
        int vfo_freq = settings.value(p + "frequency").toInt() + mix_offset;
        int data_rate = settings.value(p + "data_rate").toInt();
        int out_rate = settings.value(p + "out_rate").toInt();
        if (out_rate == 0 && data_rate > 0) {
            switch (data_rate) {
            case 600: out_rate = 12000; break;
            case 1200: out_rate = 24000; break;
            default: out_rate = 48000; break;
            }
        }
        int filterbw = settings.value(p + "filter_bandwidth").toInt();
        int main_vfo_freq = 0, main_vfo_out_rate = Fs, main_idx = 0;
        for (int a = 0; a < VFOmain.length(); a++) {
            int diff = std::abs((center_frequency - VFOmain.at(a)->getMixerFreq()) - vfo_freq);
            if (diff < VFOmain.at(a)->getOutRate() && !VFOmain.at(a)->getDemodUSB()) {
                main_idx = a;
                main_vfo_freq = VFOmain.at(a)->getMixerFreq();
                main_vfo_out_rate = VFOmain.at(a)->getOutRate();
                break;
            }
        }
        pVFO->setZmqTopic(settings.value(p + "topic"));
        pVFO->setZmqAddress(zmq_address);
        int lateDecimate = 0;
        if ((main_vfo_out_rate / 48000) == 5) {
            pVFO->setDecimationCount(int(log2(main_vfo_out_rate / (5 * out_rate))));
            lateDecimate = 5;
        } else if ((main_vfo_out_rate / 48000) == 6) {
            pVFO->setDecimationCount(int(log2(main_vfo_out_rate / (6 * out_rate))));
            lateDecimate = 6;
        } else {
            pVFO->setDecimationCount(int(log2(Fs / out_rate)) - int(log2(Fs / main_vfo_out_rate)));
        }
        pVFO->setFilterBandwidth(filterbw);
        pVFO->setGain((float)settings.value(p + "gain").toFloat() / 100);
        pVFO->setMixerFreq((center_frequency - main_vfo_freq) - vfo_freq);
        pVFO->setFs(main_vfo_out_rate);
        pVFO->setCompressonStyle(1);
        pVFO->init(main_vfo_out_rate / bufsplit, true, lateDecimate);
        VFOsub[main_idx].push_back(pVFO);
        topics.push_back(settings.value(p + "topic").toStdString());
    }
    sdrj *radio = new sdrj(0);
    radio->setVFOs(&VFOmain);
    radio->setDCCorrection(dc);
    radio->fftVFOSlot("none");   // sdrj::emitFFT is otherwise uninitialised (sdrj.cpp:4-18)
    if (fft_sel) {
        radio->fftVFOSlot(fft_sel);
        for (int m = 0; m < 3; m++)
            for (int a = 0; a < VFOsub[m].length(); a++) VFOsub[m].at(a)->fftVFOSlot(fft_sel);
    }

    // ---- input ----
    FILE *fi = fopen(in_path, "rb");
    if (!fi) { fprintf(stderr, "cannot read %s\n", in_path); return 1; }
    fseek(fi, 0, SEEK_END);
    long nbytes = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    std::vector<unsigned char> iq(nbytes);
    if (fread(iq.data(), 1, nbytes, fi) != (size_t)nbytes) { fprintf(stderr, "short read\n"); return 1; }
    fclose(fi);
    long nblocks = nbytes / buflen;
    if (max_blocks >= 0 && max_blocks < nblocks) nblocks = max_blocks;

    std::map<std::string, FILE *> pcm;
    std::vector<FILE *> mtap;
    FILE *frames = 0;
    if (!timing) {
        std::string d = out_dir;
        frames = fopen((d + "/frames.txt").c_str(), "w");
        if (!frames) { fprintf(stderr, "cannot write into %s\n", out_dir); return 1; }
        if (main_tap)
            for (int k = 0; k < VFOmain.length(); k++)
                mtap.push_back(fopen((d + "/main" + std::to_string(k) + ".cf32").c_str(), "wb"));
    }
    g_capture = !timing;
    if (fft_sel && !timing) {
        g_fft_out = fopen((std::string(out_dir) + "/fft.cf32").c_str(), "wb");
        g_fft_log = fopen((std::string(out_dir) + "/fft.txt").c_str(), "w");
    }

    std::vector<float> fl(buflen);
    double seconds = 0;
    // everything the callbacks published since the last call goes to the files
    auto flush_outputs = [&]() {
        for (const ZmqMessage &m : g_messages) {
            std::string topic = m.parts.size() > 0 ? m.parts[0] : "";
            unsigned rate = 0;
            if (m.parts.size() > 1 && m.parts[1].size() == 4) memcpy(&rate, m.parts[1].data(), 4);
            size_t plen = m.parts.size() > 2 ? m.parts[2].size() : 0;
            for (unsigned char c : topic) fprintf(frames, "%02x", c);
            fprintf(frames, " %u %zu %zu\n", rate, plen, m.parts.size());
            std::string name(topic.c_str());   // stop at an embedded NUL, if any
            FILE *&fp = pcm[name];
            if (!fp) fp = fopen((std::string(out_dir) + "/" + name + ".pcm").c_str(), "wb");
            if (fp && plen) fwrite(m.parts[2].data(), 1, plen, fp);
        }
        g_messages.clear();
    };
    if (via.compare(0, 8, "callback") == 0 && !timing) {
        // librtlsdr's thread and the dispatcher thread, played in turn by this one thread
        const int burst = via.size() > 9 ? atoi(via.c_str() + 9) : 1;
        g_have_dongle = true;
        g_audio_sink = radio;
        if (!radio->OpenRtl(0)) { fprintf(stderr, "OpenRtl failed\n"); return 1; }
        radio->StartRtl(Fs, center_frequency, buflen);             // QtConcurrent::run is inert in the shim: no threads start
        long b = 0, delivered = 0;
        while (b < nblocks) {
            for (int k = 0; k < burst && b < nblocks; k++, b++)
                radio->rtlsdr_callback(iq.data() + b * (long)buflen, (uint32_t)buflen);
            try { radio->demod_dispatcher(); } catch (const QShimWouldBlock &) {}
            delivered += 0;
            flush_outputs();
        }
        FILE *st = fopen((std::string(out_dir) + "/ingest.txt").c_str(), "w");
        fprintf(st, "callbacks %ld burst %d\n", nblocks, burst);
        fclose(st);
    } else if (via.compare(0, 6, "rtltcp") == 0 && !timing) {
        std::vector<long> sizes;
        for (const char *q = via.c_str() + (via.size() > 7 ? 7 : via.size()); *q;) {
            char *e = 0;
            long v = strtol(q, &e, 10);
            if (e == q) break;
            if (v > 0) sizes.push_back(v);
            q = *e ? e + 1 : e;
        }
        if (sizes.empty()) sizes.push_back(65536);
        qshim_tcp_connect_ok = true;
        int gain = 14;
        if (!radio->start_tcp_rtl("127.0.0.1:1234", Fs, center_frequency, gain)) { fprintf(stderr, "start_tcp_rtl failed\n"); return 1; }
        QTcpSocket *sock = qshim_last_socket;
        FILE *tx = fopen((std::string(out_dir) + "/tcp_tx.bin").c_str(), "wb");
        fwrite(sock->tx.data(), 1, sock->tx.size(), tx);
        fclose(tx);
        // rtl_tcp greets with "RTL0" + tuner type + gain count (12 bytes), delivered on its own
        const unsigned char hello[12] = {'R', 'T', 'L', '0', 0, 0, 0, 5, 0, 0, 0, 29};
        sock->rx.append((const char *)hello, 12);
        radio->readyRead();
        long at = 0, total = nblocks * (long)buflen, k = 0, calls = 0;
        while (at < total) {
            long n = sizes[(size_t)(k++ % (long)sizes.size())];
            if (n > total - at) n = total - at;
            sock->rx.append((const char *)iq.data() + at, (size_t)n);
            at += n;
            radio->readyRead();                                       // one signal per arrival
            calls++;
            flush_outputs();
        }
        FILE *st = fopen((std::string(out_dir) + "/ingest.txt").c_str(), "w");
        fprintf(st, "arrivals %ld left_in_socket %zu block_bytes %d\n", calls, sock->rx.size(), (int)radio->tcpFloats.size());
        fclose(st);
    } else
    for (long b = 0; b < nblocks; b++) {
        const unsigned char *src = iq.data() + b * (long)buflen;
        g_callback = b;
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < buflen; i++) fl[i] = radio->floats.at(src[i]);   // sdr.cpp:122-129
        radio->demodData(fl.data(), buflen);
        auto t1 = std::chrono::steady_clock::now();
        if (b >= skip_blocks) seconds += std::chrono::duration<double>(t1 - t0).count();
        if (timing) continue;
        flush_outputs();
        for (size_t k = 0; k < mtap.size(); k++) {
            vfo *mv = VFOmain.at((int)k);
            // getOutRate() == Fs / 2^decimateCount  (vfo.cpp:217-222)
            int dcount = (int)lround(log2((double)Fs / mv->getOutRate()));
            const std::vector<cpx_typef> &v = mv->decimate[dcount];
            fwrite(v.data(), sizeof(cpx_typef), v.size(), mtap[k]);
        }
    }
    if (timing) {
        printf("%ld %.6f\n", (nblocks - skip_blocks) * (long)(buflen / 2), seconds);
        return 0;
    }
    for (auto &kv : pcm) if (kv.second) fclose(kv.second);
    for (FILE *f : mtap) if (f) fclose(f);
    if (g_fft_out) { fclose(g_fft_out); fclose(g_fft_log); }
    fclose(frames);
    return 0;
}
