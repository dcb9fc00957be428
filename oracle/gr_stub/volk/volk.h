// VOLK 2 generic kernels (the *_generic implementations: plain loops over libm / IEEE operators), for the
// handful of calls cessb/clipper_cc_impl.cc and cessb/stretcher_cc_impl.cc make.  TEST INFRASTRUCTURE.
#pragma once
#include <complex>
#include <cmath>
#include <cstddef>
typedef std::complex<float> lv_32fc_t;
static inline size_t volk_get_alignment() { return 32; }
static inline void volk_32fc_magnitude_32f(float* m, const lv_32fc_t* c, unsigned n) { for (unsigned i = 0; i < n; i++) m[i] = sqrtf(c[i].real() * c[i].real() + c[i].imag() * c[i].imag()); }
static inline void volk_32f_x2_min_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] < b[i] ? a[i] : b[i]; }
static inline void volk_32f_x2_max_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] > b[i] ? a[i] : b[i]; }
static inline void volk_32f_cos_32f(float* b, const float* a, unsigned n) { for (unsigned i = 0; i < n; i++) b[i] = cosf(a[i]); }
static inline void volk_32f_sin_32f(float* b, const float* a, unsigned n) { for (unsigned i = 0; i < n; i++) b[i] = sinf(a[i]); }
static inline void volk_32f_x2_multiply_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] * b[i]; }
static inline void volk_32f_x2_add_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] + b[i]; }
static inline void volk_32f_x2_subtract_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] - b[i]; }
static inline void volk_32f_x2_divide_32f(float* c, const float* a, const float* b, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] / b[i]; }
static inline void volk_32f_s32f_multiply_32f(float* c, const float* a, float s, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = a[i] * s; }
static inline void volk_32f_x2_interleave_32fc(lv_32fc_t* c, const float* re, const float* im, unsigned n) { for (unsigned i = 0; i < n; i++) c[i] = lv_32fc_t(re[i], im[i]); }
static inline void volk_32fc_deinterleave_real_32f(float* re, const lv_32fc_t* c, unsigned n) { for (unsigned i = 0; i < n; i++) re[i] = c[i].real(); }
static inline void volk_32fc_deinterleave_imag_32f(float* im, const lv_32fc_t* c, unsigned n) { for (unsigned i = 0; i < n; i++) im[i] = c[i].imag(); }

// VOLK 2.x volk_32fc_s32f_power_spectrum_32f_generic (the _a entry point dispatches to it on x86 builds without libsimdmath)
#include <cstdlib>
static inline void* volk_malloc(size_t n, size_t al) { void* p = nullptr; return posix_memalign(&p, al < sizeof(void*) ? sizeof(void*) : al, n ? n : al) == 0 ? p : nullptr; }
static inline void volk_free(void* p) { free(p); }
static inline float volk_log2f_non_ieee(float f) { const float r = log2f(f); return std::isinf(r) ? copysignf(127.0f, r) : r; }
static inline void volk_32fc_s32f_power_spectrum_32f_a(float* logPower, const lv_32fc_t* in, const float norm, unsigned n)
{
    const float* p = reinterpret_cast<const float*>(in);
    const float inorm = 1.0f / norm;
    for (unsigned i = 0; i < n; i++) {
        const float re = p[2 * i] * inorm, im = p[2 * i + 1] * inorm;
        logPower[i] = 3.01029995663981209120f * volk_log2f_non_ieee(re * re + im * im);
    }
}
