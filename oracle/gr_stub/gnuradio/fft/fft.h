// stand-in for gnuradio/fft/fft.h (oracle/_ref build of rx_fft.cpp only): same buffers-and-execute shape as gr::fft::fft_complex_fwd;
// the transform itself is the oracle's definition (qo_dft_forward in libqrl_oracle.so: radix-2 in double, rounded to float once) --
// FFTW is not available offline, so this build pins rx_fft.cpp's buffering / windowing / drop / shift logic, not FFTW's rounding.
#pragma once
#include <complex>
#include <vector>
extern "C" void qo_dft_forward(const float* in_c, float* out_c, int n);
namespace gr { namespace fft {
namespace window { enum win_type { WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2, WIN_RECTANGULAR = 3, WIN_KAISER = 4, WIN_BLACKMAN_hARRIS = 5,
                                   WIN_BLACKMAN_HARRIS = 5, WIN_BARTLETT = 6, WIN_FLATTOP = 7 }; }
class fft_complex_fwd {
public:
    explicit fft_complex_fwd(int n, int = 1) : d_in(n), d_out(n) {}
    std::complex<float>* get_inbuf() { return d_in.data(); }
    std::complex<float>* get_outbuf() { return d_out.data(); }
    void execute() { qo_dft_forward(reinterpret_cast<const float*>(d_in.data()), reinterpret_cast<float*>(d_out.data()), static_cast<int>(d_in.size())); }
private:
    std::vector<std::complex<float>> d_in, d_out;
};
}}
