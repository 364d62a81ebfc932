// Drop-in for gnuradio/firfilter.h (firfilter.h:6-54): the firdes window designer.
// low_pass() keeps its signature (QVector<float> -> std::vector<float>) and throws
// std::out_of_range where the reference does (firfilter.cpp:122-134). filter()/setTaps()
// are dead code in the reference (never called); they are kept and run through sdrb_fir_ex.
#ifndef FIRFILTER_H
#define FIRFILTER_H
#include "sdrb_types.h"

class firfilter {
public:
    enum win_type {
        WIN_NONE = -1, WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2, WIN_RECTANGULAR = 3, WIN_KAISER = 4,
        WIN_BLACKMAN_hARRIS = 5, WIN_BLACKMAN_HARRIS = 5, WIN_BARTLETT = 6, WIN_FLATTOP = 7,
    };
    firfilter();
    ~firfilter();
    double filter(double in);
    void setTaps(std::vector<double> taps);
    std::vector<float> low_pass(double gain, double sampling_freq, double cutoff_freq, double transition_width,
                                win_type window_type, double beta);

private:
    class FIR *impl;
};
#endif
