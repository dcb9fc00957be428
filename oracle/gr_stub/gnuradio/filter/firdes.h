// gr::filter::firdes::root_raised_cosine forwards to the oracle's restatement.
#pragma once
#include <vector>
extern "C" int qo_firdes_rrc(double gain, double fs, double symrate, double alpha, int ntaps, float* out, int cap);
namespace gr { namespace filter { struct firdes {
    static std::vector<float> root_raised_cosine(double gain, double fs, double symrate, double alpha, int ntaps)
    {
        std::vector<float> t(ntaps + 2);
        const int n = qo_firdes_rrc(gain, fs, symrate, alpha, ntaps, t.data(), ntaps + 2);
        t.resize(n > 0 ? n : 0);
        return t;
    }
}; }}
