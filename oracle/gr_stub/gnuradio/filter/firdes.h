// gr::filter::firdes::root_raised_cosine forwards to the oracle's restatement.
#pragma once
#include <vector>
#include <gnuradio/fft/fft.h>
extern "C" void qo_window_build(int win, int ntaps, float* w);
extern "C" int qo_firdes_rrc(double gain, double fs, double symrate, double alpha, int ntaps, float* out, int cap);
namespace gr { namespace filter { struct firdes {
    static std::vector<float> window(gr::fft::window::win_type type, int ntaps, double /*beta*/)
    {
        std::vector<float> w(ntaps);
        qo_window_build(static_cast<int>(type), ntaps, w.data());
        return w;
    }
    static std::vector<float> root_raised_cosine(double gain, double fs, double symrate, double alpha, int ntaps)
    {
        std::vector<float> t(ntaps + 2);
        const int n = qo_firdes_rrc(gain, fs, symrate, alpha, ntaps, t.data(), ntaps + 2);
        t.resize(n > 0 ? n : 0);
        return t;
    }
}; }}
