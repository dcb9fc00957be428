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
