// C-ABI shim over the ONE file of the reference's hot path that compiles without GNU Radio/Qt:
// /root/reference/src/gr/emphasis.cpp (gr::calculate_deemph_taps / calculate_preemph_taps).
// Test infrastructure only: lets tests pin oracle/qrl_oracle.c:qo_deemph_taps against the real reference.
#include "emphasis.h"
extern "C" void ref_deemph_taps(int fs, double tau, double* a2, double* b2)
{
    std::vector<double> a, b;
    gr::calculate_deemph_taps(fs, tau, a, b);
    a2[0] = a[0]; a2[1] = a[1]; b2[0] = b[0]; b2[1] = b[1];
}
extern "C" void ref_preemph_taps(int fs, double tau, double fh, double* a2, double* b2)
{
    std::vector<double> a, b;
    gr::calculate_preemph_taps(fs, tau, a, b, fh);
    a2[0] = a[0]; a2[1] = a[1]; b2[0] = b[0]; b2[1] = b[1];
}
