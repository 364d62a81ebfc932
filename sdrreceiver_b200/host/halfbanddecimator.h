// Drop-in for halfbanddecimator.h (halfbanddecimator.h:12-26, .cpp:43-72).
// decimate() runs one callback block on the GPU (sdrb_halfband) and carries the queue head with
// the reference's off-by-one (dsp.cpp:163-173). taps = 11 (what vfo.cpp passes, vfo.cpp:130), 23
// and 51 filter; other lengths output zeros exactly like the reference, whose constructor and
// FIRUpdateAndProcessHalfBandQueue have no case for them (15 and 21 have unused tables).
#ifndef HALFBANDDECIMATOR_H
#define HALFBANDDECIMATOR_H
#include "sdrb_types.h"

class HalfBandDecimator {
public:
    HalfBandDecimator(int taps, int inlen);
    ~HalfBandDecimator();
    void decimate(const std::vector<cpx_typef> &in, std::vector<cpx_typef> &out);

private:
    float *d_in, *d_out, *d_hist;
    int cap, ntaps;
};
#endif
