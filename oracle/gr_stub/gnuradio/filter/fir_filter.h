// gr::filter::kernel::fir_filter_{ccf,ccc}: taps stored reversed, filter(in) = sum_k taps[k] * in[ntaps-1-k]
// (GNU Radio 3.10 gr-filter; VOLK dot product restated as a plain sequential loop).  TEST INFRASTRUCTURE.
#pragma once
#include "../block.h"
namespace gr { namespace filter { namespace kernel {
template <class TAP>
class fir_filter_c {
public:
    explicit fir_filter_c(const std::vector<TAP>& taps) : d_taps(taps.rbegin(), taps.rend()) {}
    gr_complex filter(const gr_complex* in) const
    {
        gr_complex acc(0.0f, 0.0f);
        for (size_t k = 0; k < d_taps.size(); k++) acc += in[k] * d_taps[k];
        return acc;
    }
    unsigned ntaps() const { return static_cast<unsigned>(d_taps.size()); }
private:
    std::vector<TAP> d_taps;
};
typedef fir_filter_c<float> fir_filter_ccf;
typedef fir_filter_c<gr_complex> fir_filter_ccc;
}}}
