// Drop-in for vfo.h (vfo.h:11-116). Same setters/getters, init(), process(), public
// decimate[] and mpVFOs; QString -> std::string, QVector -> std::vector, no QObject.
//
// A vfo is a node of the two-level tree MainWindow builds (mainwindow.cpp:98-235). process()
// is called on tree roots (main VFOs); the root compiles itself and its children into one GPU
// plan on first use (sdrb_plan_create) and from then on every call runs the whole subtree on
// the device: mix + half-band cascade, then per child mix, cascade, /late FIR, USB demod,
// low-pass, gain, int16 and the ZMQ publish. Calling process() on a child throws: the
// reference only ever reaches children through their parent (vfo.cpp:253-266). A root without
// children is the reference's IQ forwarder: mix + cascade, vfo::compress (vfo.cpp:389-424) and,
// when it has a topic, one ZMQ message of packed IQ bytes per callback.
#ifndef VFO_H
#define VFO_H
#include <cstdint>
#include <functional>
#include "sdrb_types.h"
#include "zmqpublisher.h"
#include "halfbanddecimator.h"
#include "oscillator.h"

struct sdrb_plan;
struct sdrb_bank;

class vfo {
public:
    ~vfo();
    vfo(void *parent = 0);

    void init(int samplesPerBuffer, bool bind, int lateDecimate = 0);
    void process(const std::vector<cpx_typef> &samples);
    void setZmqAddress(std::string bind);
    void setZmqTopic(std::string topic);
    void setScaleComp(int scale);
    void setFs(int samplerate);
    void setDecimationCount(int count);
    void setMixerFreq(double freq);
    double getMixerFreq();
    int getOutRate();
    void setOffsetBandwidth(double bw);
    void setFilterBandwidth(double bw);
    void setGain(float g);
    void setDemodUSB(bool usb);
    bool getDemodUSB();
    void setCompressonStyle(int st);
    void setFilter(bool filter, int bw = 0);
    void setVFOs(std::vector<vfo *> *pVFOs);
    std::vector<cpx_typef> decimate[9];
    std::vector<vfo *> *mpVFOs;

    // signal fftData(const std::vector<cpx_typef>&) of the reference, as a callback
    std::function<void(const std::vector<cpx_typef> &)> fftData;
    void fftVFOSlot(std::string topic);

    // what the last callback produced for this (leaf) VFO: the ZMQ payload
    const std::vector<short> &lastAudio() const { return transmit_usb; }
    // ... and for a childless root: the vfo::compress payload (transmit_iq, vfo.h:69)
    const std::vector<signed char> &lastForward() const { return transmit_iq; }
    uint32_t getOutputRate() const { return outputRate; }
    int getSamplesPerBuffer() const { return samplesPerBuffer; }
    const std::string &getZmqTopic() const { return zmqTopic; }

private:
    friend class sdrj;
    std::string zmqAddress, zmqTopic;
    int Fs;
    bool zmqBind;
    static ZmqPublisher bind_publisher;
    ZmqPublisher connect_publisher;
    std::vector<short> transmit_usb;
    std::vector<signed char> transmit_iq;
    int decimateCount;
    uint32_t outputRate;
    float gain;
    double mixer_freq;
    bool demodUSB, filterAudio;
    int cstyle, filterbw, offsetbw, scalecomp;
    int samplesPerBuffer, lateDecimate;
    bool emitFFT;
    void transmitData();
    void emit_selected(sdrb_bank *root_bank, int sub_idx);
    // GPU side (roots only)
    sdrb_plan *plan;
    sdrb_bank *bank;
    std::vector<short> pcm_record;
    void compile_tree();
};
#endif
