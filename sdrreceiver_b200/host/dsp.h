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

// DelayThing<T> (jonti/dsp.h:79-126): an integer delay line. Same public interface and behaviour -- a line of length + 1
// slots, a sample comes back `length` calls after it went in, setLength() keeps what the line held (the reference
// resize()s), findmaxpos() reports the first largest sample counted from the oldest -- written around one push() helper.
// Pure data movement: it stays a host template (the kernels realise the 62-sample delay of vfo.cpp:136 as an index offset).
template <class T>
class DelayThing {
public:
    DelayThing() : head(0) { setLength(12); }
    void setLength(int length) {
        assert(length + 1 > 0);
        line.resize((size_t)length + 1);
        head = 0;
    }
    void update(T &data) { data = push(data); }
    T update_dont_touch(T data) { return push(data); }
    int findmaxpos(T &maxval) {
        const size_t n = line.size();
        size_t best = 0;
        maxval = line[head];
        for (size_t i = 1; i < n; ++i) {
            const T &v = line[(head + i) % n];
            if (v > maxval) { maxval = v; best = i; }
        }
        return (int)best;
    }

private:
    T push(const T &v) {                                   // store, advance, hand back the oldest sample
        line[head] = v;
        head = (head + 1) % line.size();
        return line[head];
    }
    std::vector<T> line;
    size_t head;
};
#endif
