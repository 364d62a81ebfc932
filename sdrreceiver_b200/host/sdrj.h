// Drop-in for the ingest half of sdrj.h / jonti/sdr.h (sdrj.h:18-75, sdr.h:54-73):
// setVFOs, setDCCorrection, fftVFOSlot, demodData(const float*, int) and the raw callback
// rtlsdr_callback(unsigned char*, uint32_t) with the `floats` table. Device I/O (librtlsdr,
// rtl_tcp) is out of scope; whoever owns the dongle calls one of the
// two entry points per callback buffer, exactly like the reference's dispatcher does.
//
// The whole VFO tree (all main VFOs and their children) is compiled into one GPU plan on the
// first buffer; each call then runs uint8 -> DC removal -> every VFO -> int16 on the device
// and publishes one ZMQ message per leaf VFO.
#ifndef SDRJ_H
#define SDRJ_H
#include "vfo.h"

class sdrj {
public:
    explicit sdrj(void *parent = 0);
    ~sdrj();
    void setDCCorrection(bool correct);
    void setVFOs(std::vector<vfo *> *pVFOs);
    void fftVFOSlot(std::string topic);
    void demodData(const float *data, int len);                    // sdrj.cpp:266-305
    void rtlsdr_callback(unsigned char *buf, uint32_t len);        // jonti/sdr.cpp:100-145 + dispatcher
    std::vector<float> floats;                                     // (i - 127), sdr.cpp:43-49
    std::function<void(const std::vector<cpx_typef> &)> fftData;   // signal of the reference
    bool publishEnabled;                                           // extension: tests switch ZMQ off

private:
    void compile_tree(int block);
    void run(const unsigned char *bytes, uint32_t len);
    std::vector<vfo *> *mpVFOs;
    std::vector<vfo *> leaves;
    bool correctDC, emitFFT;
    int count;
    sdrb_plan *plan;
    sdrb_bank *bank;
    std::vector<unsigned char> staging;
    std::vector<short> pcm_record;
    std::vector<cpx_typef> samples;
};
#endif
