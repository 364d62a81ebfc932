// Drop-in for the reference's oscillator.h (oscillator.h:12-29, oscillator.cpp:4-50).
// Same constructor, tick() and public _vector. The 1-second table comes from
// sdrb_nco_table() (bit-identical recursion); tick() is a table walk, as in the reference.
#ifndef OSCILLATOR_H
#define OSCILLATOR_H
#include "sdrb_types.h"

class Oscillator {
public:
    Oscillator(double sampleRate, double Frequency);
    void tick();
    cpx_typef _vector;
    ~Oscillator();
    // extension: the table on the device, for sdrb_nco_mix (what vfo::process does with it)
    const float *deviceTable();
    int tableLength() const { return length; }

private:
    std::vector<cpx_typef> queue;
    int queuePtr;
    int length;
    float *d_table;
};
#endif
