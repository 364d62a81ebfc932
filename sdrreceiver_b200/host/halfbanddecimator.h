// Drop-in for halfbanddecimator.h (halfbanddecimator.h:12-26, .cpp:43-72).
// decimate() runs one callback block on the GPU (sdrb_halfband11) and carries the 11-sample
// queue head with the reference's off-by-one (dsp.cpp:163-173). Only taps == 11 exists on the
// GPU (the only value vfo.cpp passes, vfo.cpp:130); other values throw.
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
    int cap;
};
#endif
