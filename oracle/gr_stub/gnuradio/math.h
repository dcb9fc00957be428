// gr::fast_atan2f lives in gnuradio-runtime (not under /root/reference); the stand-in forwards to the oracle's
// restatement of it (table + octant fix-up), which is what the pin then holds fixed.
#pragma once
#include "block.h"
extern "C" float qo_fast_atan2f(float y, float x);
namespace gr {
static inline float fast_atan2f(float y, float x) { return qo_fast_atan2f(y, x); }
static inline float fast_atan2f(gr_complex z) { return qo_fast_atan2f(z.imag(), z.real()); }
}
