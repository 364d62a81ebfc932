// Drop-in for jonti/dsp.h (dsp.h:34-126): FIR, FIRHilbert, DelayThing<T>.
// The per-sample calls keep their signatures and semantics; each one runs as a (tiny) batch
// on the GPU through sdrb_fir_ex. Code that wants throughput uses the block calls
// (process()), which is what the vfo facade does internally.
#ifndef DSP_F_H
#define DSP_F_H
#include <cassert>
#include "sdrb_types.h"

class FIR {
public:
    FIR(int _NumberOfPoints, int queuesz);
    ~FIR();
    float FIRUpdateAndProcess(float sig);                 // dsp.cpp:59-71 (newest sample excluded)
    void FIRUpdate(float sig);                            // dsp.cpp:150-154
    void FIRSetPoint(int point, float value);             // dsp.cpp:177-182
    // half-band queue entry points (dsp.cpp:96-173); 11 taps only, like every caller in vfo.cpp
    float FIRUpdateAndProcessHalfBandQueue(float sig);
    void FIRUpdateQueue(float sig);
    void FIRQueueBackToFront();
    // block form: out[m] = sum_i points[i]*x[decim*m - N + i], state carried between calls
    // include_newest = true gives y[m] = sum_i points[i]*x[decim*m - N + 1 + i] (FIRHilbert / firfilter form)
    void process(const float *in, int n, float *out, int decim = 1, bool include_newest = false);

    float *points;
    int NumberOfPoints;
    float outsum;

private:
    void sync_taps();
    std::vector<float> pending;     // FIRUpdate()d samples not yet pushed to the device
    std::vector<float> hbq;         // half-band queue of the current block (host staging)
    float *d_taps, *d_hist, *d_io;
    float hb_hist[11];
    int io_cap;
    bool taps_dirty;
};

class FIRHilbert {
public:
    FIRHilbert(int len, int Fs);
    ~FIRHilbert();
    double FIRUpdateAndProcess(float sig);                // dsp.cpp:218-231 (newest sample included)
    void process(const float *in, int n, float *out);     // block form
    float *points;
    int NumberOfPoints;
    float outsum;

private:
    float *d_taps, *d_hist, *d_io;
    int io_cap;
};

template <class T>
class DelayThing {                                        // dsp.h:79-126: pure data movement, stays on the host
public:
    DelayThing() { setLength(12); }
    void setLength(int length) {
        length++;
        assert(length > 0);
        buffer.assign((size_t)length, T());
        buffer_ptr = 0;
        buffer_sz = (int)buffer.size();
    }
    void update(T &data) {
        buffer[(size_t)buffer_ptr] = data;
        buffer_ptr++; buffer_ptr %= buffer_sz;
        data = buffer[(size_t)buffer_ptr];
    }
    T update_dont_touch(T data) {
        buffer[(size_t)buffer_ptr] = data;
        buffer_ptr++; buffer_ptr %= buffer_sz;
        return buffer.at((size_t)buffer_ptr);
    }
    int findmaxpos(T &maxval) {
        int maxpos = 0;
        maxval = buffer[(size_t)buffer_ptr];
        for (int i = 0; i < buffer_sz; i++) {
            if (buffer[(size_t)buffer_ptr] > maxval) { maxval = buffer[(size_t)buffer_ptr]; maxpos = i; }
            buffer_ptr++; buffer_ptr %= buffer_sz;
        }
        return maxpos;
    }

private:
    std::vector<T> buffer;
    int buffer_ptr;
    int buffer_sz;
};
#endif
