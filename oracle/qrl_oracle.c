/*
 * qrl_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see qrl_oracle.h).
 *
 * Plain-C restatement of the GNU Radio 3.10 blocks the reference wires up in
 *   /root/reference/src/gr/gr_demod_{nbfm,4fsk,qpsk}.cpp, gr_mod_{4fsk,qpsk}.cpp
 * with the literal parameters of gr_demod_base.cpp:203-228 / gr_mod_base.cpp:155-180.
 * Block semantics follow SURVEY.md Appendix A (GNU Radio / VOLK are not vendored by the
 * reference and are absent here): PARITY UNPINNED against real GNU Radio.
 *
 * Numerics contract shared with the CUDA path (bit-exactness by construction):
 *   - compiled with -ffp-contract=off; every fused multiply-add is an explicit fmaf();
 *   - FIR dot products use ONE fixed order (fir_dot below): polyphase branches r = j mod D are
 *     accumulated oldest-sample-first with fmaf into 32 "lane" partial sums (lane = r mod 32),
 *     which are then combined by a fixed butterfly tree (offsets 16,8,4,2,1).  VOLK's real order
 *     depends on the host SIMD width, so GNU Radio itself has no single bit-exact result; the
 *     sequential order is kept (qo_set_fir_order(1)) so tests can bound the difference;
 *   - sin/cos is qo_sincosf (Cody-Waite + Cephes polynomials in explicit fmaf steps), not libm;
 *   - tables (atan, tanh, MMSE interpolator, fxpt sine) are generated here from closed forms.
 */
#define _GNU_SOURCE
#include "qrl_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int g_fir_order = 0;
static int g_fm_literal = 0;
static int g_dbg_bypass_fll = 0;
void qo_dbg_bypass_fll(int on) { g_dbg_bypass_fll = on; }
void qo_set_fir_order(int o) { g_fir_order = o; }
void qo_set_fm_literal(int on) { g_fm_literal = on; }

/* ------------------------------------------------------------------ growable vectors */
typedef struct { unsigned char* d; size_t n, cap, isz; } qvec;
static void qv_init(qvec* v, size_t isz) { v->d = NULL; v->n = 0; v->cap = 0; v->isz = isz; }
static void qv_free(qvec* v) { free(v->d); v->d = NULL; v->n = v->cap = 0; }
static void* qv_grow(qvec* v, size_t extra)
{
    if (v->n + extra > v->cap) {
        size_t nc = v->cap ? v->cap : 1024;
        while (nc < v->n + extra) nc *= 2;
        v->d = (unsigned char*)realloc(v->d, nc * v->isz);
        v->cap = nc;
    }
    return v->d + v->n * v->isz;
}
static void qv_push(qvec* v, const void* items, size_t n)
{
    if (!n) return;
    void* p = qv_grow(v, n);
    memcpy(p, items, n * v->isz);
    v->n += n;
}
static void qv_push_zero(qvec* v, size_t n)
{
    if (!n) return;
    void* p = qv_grow(v, n);
    memset(p, 0, n * v->isz);
    v->n += n;
}
static void qv_drop(qvec* v, size_t n)
{
    if (!n) return;
    memmove(v->d, v->d + n * v->isz, (v->n - n) * v->isz);
    v->n -= n;
}
static inline void qv_pushf(qvec* v, float x) { *(float*)qv_grow(v, 1) = x; v->n++; }
static inline void qv_pushc(qvec* v, float re, float im) { float* p = (float*)qv_grow(v, 1); p[0] = re; p[1] = im; v->n++; }
static inline void qv_pushb(qvec* v, unsigned char b) { *(unsigned char*)qv_grow(v, 1) = b; v->n++; }

/* ------------------------------------------------------------------ design: windows / firdes (Appendix A1) */
static double win_max_att(int win)
{
    switch (win) {
    case QO_WIN_HAMMING: return 53; case QO_WIN_HANN: return 44; case QO_WIN_BLACKMAN: return 74;
    case QO_WIN_RECT: return 21; case QO_WIN_BLACKMAN_HARRIS: return 92; default: return 53;
    }
}
/* gr::fft::window::build: cosine windows evaluated in float (coswindow) */
static void win_build(int win, int ntaps, float* w)
{
    float M = (float)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) {
        switch (win) {
        case QO_WIN_HAMMING:
            w[n] = (float)(0.54 - 0.46 * cos((2 * M_PI * n) / M));
            break;
        case QO_WIN_HANN:
            w[n] = (float)(0.5 - 0.5 * cos((2 * M_PI * n) / M));
            break;
        case QO_WIN_BLACKMAN: {
            float c0 = 0.42f, c1 = 0.5f, c2 = 0.08f;
            w[n] = c0 - c1 * cosf((float)((2.0 * M_PI * n) / M)) + c2 * cosf((float)((4.0 * M_PI * n) / M));
            break; }
        case QO_WIN_BLACKMAN_HARRIS: {
            float c0 = 0.35875f, c1 = 0.48829f, c2 = 0.14128f, c3 = 0.01168f;
            w[n] = c0 - c1 * cosf((float)((2.0 * M_PI * n) / M)) + c2 * cosf((float)((4.0 * M_PI * n) / M))
                   - c3 * cosf((float)((6.0 * M_PI * n) / M));
            break; }
        default: w[n] = 1.0f;
        }
    }
}
static int compute_ntaps(double fs, double tw, int win)
{
    double a = win_max_att(win);
    int ntaps = (int)(a * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    return ntaps;
}
static int compute_ntaps_windes(double fs, double tw, double att)
{
    int ntaps = (int)(att * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    return ntaps;
}
static int low_pass_n(double gain, double fs, double fc, int ntaps, int win, float* taps)
{
    float* w = (float*)malloc(sizeof(float) * ntaps);
    win_build(win, ntaps, w);
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * M_PI * fc / fs;
    for (int n = -M; n <= M; n++) {
        if (n == 0) taps[n + M] = (float)(fwT0 / M_PI * w[n + M]);
        else taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
    }
    double fmax = taps[0 + M];
    for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M];
    gain /= fmax;
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain);
    free(w);
    return ntaps;
}
int qo_firdes_low_pass(double gain, double fs, double fc, double tw, int win, float* out, int cap)
{
    int nt = compute_ntaps(fs, tw, win);
    if (nt > cap) return -nt;
    return low_pass_n(gain, fs, fc, nt, win, out);
}
int qo_firdes_low_pass_2(double gain, double fs, double fc, double tw, double att, int win, float* out, int cap)
{
    int nt = compute_ntaps_windes(fs, tw, att);
    if (nt > cap) return -nt;
    return low_pass_n(gain, fs, fc, nt, win, out);
}
static int band_pass_n(double gain, double fs, double lo, double hi, int ntaps, int win, float* taps)
{
    float* w = (float*)malloc(sizeof(float) * ntaps);
    win_build(win, ntaps, w);
    int M = (ntaps - 1) / 2;
    double fwT0 = 2 * M_PI * lo / fs, fwT1 = 2 * M_PI * hi / fs;
    for (int n = -M; n <= M; n++) {
        if (n == 0) taps[n + M] = (float)((fwT1 - fwT0) / M_PI * w[n + M]);
        else taps[n + M] = (float)((sin(n * fwT1) - sin(n * fwT0)) / (n * M_PI) * w[n + M]);
    }
    double fmax = taps[0 + M];
    for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M] * cos(n * (fwT0 + fwT1) * 0.5);
    gain /= fmax;
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain);
    free(w);
    return ntaps;
}
int qo_firdes_band_pass(double gain, double fs, double lo, double hi, double tw, int win, float* out, int cap)
{
    int nt = compute_ntaps(fs, tw, win);
    if (nt > cap) return -nt;
    return band_pass_n(gain, fs, lo, hi, nt, win, out);
}
int qo_firdes_band_pass_2(double gain, double fs, double lo, double hi, double tw, double att, int win, float* out, int cap)
{
    int nt = compute_ntaps_windes(fs, tw, att);
    if (nt > cap) return -nt;
    return band_pass_n(gain, fs, lo, hi, nt, win, out);
}
/* firdes::complex_band_pass: low-pass prototype rotated to the band centre */
static int complex_band_pass_n(double gain, double fs, double lo, double hi, int ntaps, int win, float* out_c)
{
    float* lp = (float*)malloc(sizeof(float) * ntaps);
    low_pass_n(gain, fs, (hi - lo) / 2, ntaps, win, lp);
    float freq = (float)(M_PI * (hi + lo) / fs);
    float phase = 0;
    if (ntaps & 1) phase = -freq * (float)(ntaps >> 1);
    else phase = -freq / 2.0f * (float)((1 + 2 * ntaps) >> 1);
    for (int i = 0; i < ntaps; i++) {
        out_c[2 * i] = lp[i] * cosf(phase);
        out_c[2 * i + 1] = lp[i] * sinf(phase);
        phase += freq;
    }
    free(lp);
    return ntaps;
}
int qo_firdes_complex_band_pass(double gain, double fs, double lo, double hi, double tw, int win, float* out_c, int cap)
{
    int nt = compute_ntaps(fs, tw, win);
    if (nt > cap) return -nt;
    return complex_band_pass_n(gain, fs, lo, hi, nt, win, out_c);
}
int qo_firdes_complex_band_pass_2(double gain, double fs, double lo, double hi, double tw, double att, int win, float* out_c, int cap)
{
    int nt = compute_ntaps_windes(fs, tw, att);
    if (nt > cap) return -nt;
    return complex_band_pass_n(gain, fs, lo, hi, nt, win, out_c);
}
int qo_firdes_rrc(double gain, double fs, double symrate, double alpha, int ntaps, float* taps, int cap)
{
    ntaps |= 1;
    if (ntaps > cap) return -ntaps;
    double spb = fs / symrate;
    double scale = 0;
    for (int i = 0; i < ntaps; i++) {
        double x1, x2, x3, num, den;
        double xindx = i - ntaps / 2;
        x1 = M_PI * xindx / spb;
        x2 = 4 * alpha * xindx / spb;
        x3 = x2 * x2 - 1;
        if (fabs(x3) >= 0.000001) {
            if (i != ntaps / 2) num = cos((1 + alpha) * x1) + sin((1 - alpha) * x1) / (4 * alpha * xindx / spb);
            else num = cos((1 + alpha) * x1) + (1 - alpha) * M_PI / (4 * alpha);
            den = x3 * M_PI;
        } else {
            if (alpha == 1) { taps[i] = -1; scale += taps[i]; continue; }
            x3 = (1 - alpha) * x1;
            x2 = (1 + alpha) * x1;
            num = (sin(x2) * (1 + alpha) * M_PI - cos(x3) * ((1 - alpha) * M_PI * spb) / (4 * alpha * xindx)
                   + sin(x3) * spb * spb / (4 * alpha * xindx * xindx));
            den = -32 * M_PI * alpha * alpha * xindx / spb;
        }
        taps[i] = (float)(4 * alpha * num / den);
        scale += taps[i];
    }
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain / scale);
    return ntaps;
}
/* gr::filter::firdes::gaussian(gain, spb, bt, ntaps) restated (GNU Radio 3.10 gr-filter firdes.cc; used by
 * /root/reference/src/gr/gr_mod_gmsk.cpp:77-79) */
int qo_firdes_gaussian(double gain, double spb, double bt, int ntaps, float* taps, int cap)
{
    if (ntaps > cap) return -ntaps;
    double scale = 0;
    const double dt = 1.0 / spb;
    const double s = 1.0 / (sqrt(log(2.0)) / (2 * M_PI * bt));
    double t0 = -0.5 * ntaps;
    for (int i = 0; i < ntaps; i++) {
        t0++;
        const double ts = s * dt * t0;
        taps[i] = (float)exp(-0.5 * ts * ts);
        scale += taps[i];
    }
    for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] / scale * gain);
    return ntaps;
}
/* /root/reference/src/gr/emphasis.cpp:16-42 (note the float tanf inside double math) */
void qo_deemph_taps(int sample_rate, double tau, double* a, double* b)
{
    double fs = (double)sample_rate;
    double w_c = 1.0 / tau;
    double w_ca = 2.0 * fs * tanf(w_c / (2.0 * fs));
    double k = -w_ca / (2.0 * fs);
    double z1 = -1.0;
    double p1 = (1.0 + k) / (1.0 - k);
    double b0 = -k / (1.0 - k);
    b[0] = b0 * 1.0; b[1] = b0 * -z1;
    a[0] = 1.0; a[1] = -p1;
}
/* /root/reference/src/gr/emphasis.cpp:44-88 */
void qo_preemph_taps(int sample_rate, double tau, double fh, double* a, double* b)
{
    double fs = (double)sample_rate;
    if (fh <= 0.0 || fh >= fs / 2.0) fh = 0.925 * fs / 2.0;
    double w_cl = 1.0 / tau;
    double w_ch = 2.0 * M_PI * fh;
    double w_cla = 2.0 * fs * tanf(w_cl / (2.0 * fs));
    double w_cha = 2.0 * fs * tanf(w_ch / (2.0 * fs));
    double kl = -w_cla / (2.0 * fs);
    double kh = -w_cha / (2.0 * fs);
    double z1 = (1.0 + kl) / (1.0 - kl);
    double p1 = (1.0 + kh) / (1.0 - kh);
    double b0 = (1.0 - kl) / (1.0 - kh);
    double w_0dB = 2.0 * M_PI * 0.0;
    double g = fabs(1.0 - p1 * 1.0 * (cos(-w_0dB) + sin(-w_0dB))) / (b0 * fabs(1.0 - z1 * 1.0 * (cos(-w_0dB) + sin(-w_0dB))));
    b[0] = g * b0 * 1.0; b[1] = g * b0 * -z1;
    a[0] = 1.0; a[1] = -p1;
}

/* ------------------------------------------------------------------ tables */
/* gr::fast_atan2f table: 257 entries of atan(i/255) as the 7-significant-digit literals GNU Radio prints */
void qo_atan_table(float* t)
{
    char buf[64];
    for (int i = 0; i < 256; i++) {
        snprintf(buf, sizeof buf, "%e", atan((double)i / 255.0));
        t[i] = strtof(buf, NULL);
    }
    t[256] = t[255];
}
/* blocks::tanhf_lut table: tanh((i-128)/64), 256 entries */
void qo_tanh_table(float* t)
{
    for (int i = 0; i < 256; i++) t[i] = (float)tanh((double)(i - 128) / 64.0);
}
/* gr::fxpt sine table: 1024 (slope, intercept) pairs over u = (uint32 phase)>>1, f(u) = sin(u*pi/2^30) */
void qo_fxpt_sine_table(float* t)
{
    const double scale = M_PI / 1073741824.0; /* pi / 2^30 */
    const double inc = 2097152.0;             /* 2^21 */
    for (int i = 0; i < 1024; i++) {
        double a = i * inc, b = (i + 1) * inc;
        double m = (sin(b * scale) - sin(a * scale)) / (b - a);
        double c = sin(a * scale) - m * a;
        t[2 * i] = (float)m;
        t[2 * i + 1] = (float)c;
    }
}
/* MMSE 8-tap fractional interpolator bank (gr::filter::mmse_fir_interpolator, interpolator_taps.h):
 * regenerated from the MMSE criterion over |f| < 0.25 fs and rounded to the 6 significant digits the
 * upstream header prints (rows checked against the upstream values in tests/test_oracle_design.py). */
static void solve8(double A[8][9])
{
    for (int c = 0; c < 8; c++) {
        int p = c;
        for (int r = c + 1; r < 8; r++) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (p != c) for (int k = 0; k < 9; k++) { double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
        for (int r = c + 1; r < 8; r++) {
            double f = A[r][c] / A[c][c];
            for (int k = c; k < 9; k++) A[r][k] -= f * A[c][k];
        }
    }
    for (int r = 7; r >= 0; r--) {
        double s = A[r][8];
        for (int k = r + 1; k < 8; k++) s -= A[r][k] * A[k][8];
        A[r][8] = s / A[r][r];
    }
}
static double mmse_s(double d)
{
    const double B = 0.25;
    if (fabs(d) < 1e-12) return 2 * B;
    return sin(2 * M_PI * B * d) / (M_PI * d);
}
void qo_mmse_table(float* t)
{
    char buf[64];
    for (int imu = 0; imu <= 128; imu++) {
        double mu = imu / 128.0;
        double A[8][9];
        for (int a = 0; a < 8; a++) {
            for (int b = 0; b < 8; b++) A[a][b] = mmse_s((double)(a - b));
            A[a][8] = mmse_s((double)(a - 4) + mu);
        }
        solve8(A);
        for (int k = 0; k < 8; k++) {
            double v = A[k][8];
            if (fabs(v) < 1e-9) v = 0.0;
            snprintf(buf, sizeof buf, "%.5e", v);
            t[imu * 8 + k] = strtof(buf, NULL);
        }
    }
}

/* ------------------------------------------------------------------ elementary functions */
void qo_sincosf(float x, float* s, float* c)
{
    float k = rintf(x * 0.636619772f);
    float r = fmaf(k, -1.5703125f, x);
    r = fmaf(k, -4.837512969970703125e-4f, r);
    r = fmaf(k, -7.54978995489188216e-8f, r);
    int q = ((int)k) & 3;
    float z = r * r;
    float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    ps = fmaf(ps, z, -1.6666654611e-1f);
    ps = ps * z;
    float sn = fmaf(ps, r, r);
    float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    pc = fmaf(pc, z, 4.166664568298827e-2f);
    pc = pc * z;
    pc = pc * z;
    float cs = fmaf(z, -0.5f, 1.0f);
    cs = cs + pc;
    switch (q) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
    }
}
static float g_atan_tab[257];
static float g_tanh_tab[256];
static float g_mmse_tab[129 * 8];
static float g_sine_tab[2048];
static int g_tabs_ready = 0;
static void tabs_init(void)
{
    if (g_tabs_ready) return;
    qo_atan_table(g_atan_tab);
    qo_tanh_table(g_tanh_tab);
    qo_mmse_table(g_mmse_tab);
    qo_fxpt_sine_table(g_sine_tab);
    g_tabs_ready = 1;
}
/* gnuradio-runtime/lib/math/fast_atan2f.cc (Appendix A4) */
float qo_fast_atan2f(float y, float x)
{
    tabs_init();
    float x_abs, y_abs, z, alpha, angle, base_angle;
    int index;
    y_abs = fabsf(y);
    x_abs = fabsf(x);
    if (!((y_abs > 0.0f) || (x_abs > 0.0f))) return 0.0f;
    if (y_abs < x_abs) z = y_abs / x_abs; else z = x_abs / y_abs;
    if ((double)z < 0.003921569) base_angle = z;
    else {
        alpha = z * 255.0f;
        index = ((int)alpha) & 0xff;
        alpha -= (float)index;
        base_angle = g_atan_tab[index];
        base_angle += (g_atan_tab[index + 1] - g_atan_tab[index]) * alpha;
    }
    if (x_abs > y_abs) {
        if (x >= 0.0f) { if (y >= 0.0f) angle = base_angle; else angle = -base_angle; }
        else { angle = 3.14159265358979323846f; if (y >= 0.0f) angle -= base_angle; else angle = base_angle - angle; }
    } else {
        if (y >= 0.0f) { angle = 1.57079632679489661923f; if (x >= 0.0f) angle -= base_angle; else angle += base_angle; }
        else { angle = -1.57079632679489661923f; if (x >= 0.0f) angle += base_angle; else angle -= base_angle; }
    }
    return angle;
}
static inline float tanhf_lut(float x)
{
    if (x > 2.0f) return 1.0f;
    else if (x <= -2.0f) return -1.0f;
    int index = (int)(128.0f + 64.0f * x);
    if (index > 255) index = 255;
    return g_tanh_tab[index];
}
static inline float clipf(float x, float lim)
{
    if (x > lim) return lim;
    if (x < -lim) return -lim;
    return x;
}
/* digital::clock_tracking_loop::update_gains */
void qo_clock_loop_gains(float loop_bw, float damping, float ted_gain, float* alpha, float* beta)
{
    float omega_n_T = loop_bw, zeta = damping;
    float k0 = 2.0f / ted_gain;
    float k1 = expf(-zeta * omega_n_T);
    float sh = sinhf(zeta * omega_n_T);
    float cx;
    if (zeta > 1.0f) cx = coshf(omega_n_T * sqrtf(zeta * zeta - 1.0f));
    else if (zeta == 1.0f) cx = 1.0f;
    else cx = cosf(omega_n_T * sqrtf(1.0f - zeta * zeta));
    *alpha = k0 * k1 * sh;
    *beta = k0 * (1.0f - k1 * (sh + cx));
}
/* blocks::control_loop::update_gains, damping sqrt(2)/2 */
void qo_control_loop_gains(float loop_bw, float* alpha, float* beta)
{
    float damping = sqrtf(2.0f) / 2.0f;
    float denom = (float)(1.0 + 2.0 * damping * loop_bw + loop_bw * loop_bw);
    *alpha = (4 * damping * loop_bw) / denom;
    *beta = (4 * loop_bw * loop_bw) / denom;
}

/* ------------------------------------------------------------------ THE FIR dot-product order */
/* returns sum_j h[j] * x[-j*stride]  (x points at the newest sample; D = decimation of the filter) */
static float fir_dot(const float* h, int ntaps, int D, const float* x, int stride)
{
    if (g_fir_order == 1) {
        float s = 0.0f;
        for (int j = ntaps - 1; j >= 0; j--) s = fmaf(h[j], x[-(long)j * stride], s);
        return s;
    }
    float S[32];
    for (int l = 0; l < 32; l++) S[l] = 0.0f;
    for (int r = 0; r < D && r < ntaps; r++) {
        float s = S[r & 31];
        int qmax = (ntaps - 1 - r) / D;
        for (int q = qmax; q >= 0; q--) {
            int j = D * q + r;
            s = fmaf(h[j], x[-(long)j * stride], s);
        }
        S[r & 31] = s;
    }
    for (int off = 16; off >= 1; off >>= 1)
        for (int l = 0; l < off; l++) S[l] = S[l] + S[l + off];
    return S[0];
}

/* ------------------------------------------------------------------ rational resampler / FIR (A2, A3) */
typedef struct {
    int ncomp;          /* 1 = float stream, 2 = complex stream */
    int L, M;           /* interpolation, decimation */
    int nt;             /* taps per arm */
    float** arm;        /* L arms */
    float* arm_store;
    unsigned ctr;
    size_t pos;         /* index in `in` of the newest sample for the next output */
    qvec in;
    size_t hkeep;       /* history kept beyond the taps' own need (filters whose taps a run-time setter may lengthen) */
} resamp_t;

static void resamp_init(resamp_t* r, int ncomp, int L, int M, const float* taps, int ntaps)
{
    int a = L, b = M; while (b) { int t = a % b; a = b; b = t; }
    L /= a; M /= a;
    r->ncomp = ncomp; r->L = L; r->M = M;
    int padded = ((ntaps + L - 1) / L) * L;
    r->nt = padded / L;
    r->arm_store = (float*)calloc(padded, sizeof(float));
    r->arm = (float**)malloc(sizeof(float*) * L);
    for (int p = 0; p < L; p++) {
        r->arm[p] = r->arm_store + (size_t)p * r->nt;
        for (int k = 0; k < r->nt; k++) {
            int j = p + k * L;
            r->arm[p][k] = j < ntaps ? taps[j] : 0.0f;
        }
    }
    r->ctr = 0;
    r->hkeep = 0;
    qv_init(&r->in, sizeof(float) * ncomp);
    qv_push_zero(&r->in, r->nt - 1);
    r->pos = r->nt - 1;
}
static void resamp_free(resamp_t* r) { free(r->arm_store); free(r->arm); qv_free(&r->in); }
/* set_taps of a running filter (L = M = 1 users only): the new taps meet the true sample history from the next output on
 * (before the first sample: zeros, like GNU Radio's zero history) */
static void resamp_retap(resamp_t* r, const float* taps, int ntaps)
{
    free(r->arm_store); free(r->arm);
    const int L = r->L;
    int padded = ((ntaps + L - 1) / L) * L;
    r->nt = padded / L;
    r->arm_store = (float*)calloc(padded, sizeof(float));
    r->arm = (float**)malloc(sizeof(float*) * L);
    for (int p = 0; p < L; p++) {
        r->arm[p] = r->arm_store + (size_t)p * r->nt;
        for (int k = 0; k < r->nt; k++) { int j = p + k * L; r->arm[p][k] = j < ntaps ? taps[j] : 0.0f; }
    }
    if (r->pos < (size_t)(r->nt - 1)) {
        const size_t add = (size_t)(r->nt - 1) - r->pos, isz = r->in.isz, n0 = r->in.n;
        qv_push_zero(&r->in, add);
        memmove(r->in.d + add * isz, r->in.d, n0 * isz);
        memset(r->in.d, 0, add * isz);
        r->pos += add;
    }
}
static void resamp_work(resamp_t* r, const float* x, size_t n, qvec* out)
{
    qv_push(&r->in, x, n);
    const float* b = (const float*)r->in.d;
    int Drule = (r->L == 1) ? r->M : 1;
    while (r->pos < r->in.n) {
        const float* newest = b + r->pos * r->ncomp;
        const float* h = r->arm[r->ctr];
        if (r->ncomp == 2) {
            float re = fir_dot(h, r->nt, Drule, newest, 2);
            float im = fir_dot(h, r->nt, Drule, newest + 1, 2);
            qv_pushc(out, re, im);
        } else {
            qv_pushf(out, fir_dot(h, r->nt, Drule, newest, 1));
        }
        r->ctr += r->M;
        while (r->ctr >= (unsigned)r->L) { r->ctr -= r->L; r->pos++; }
    }
    const size_t need = (size_t)(r->nt - 1) > r->hkeep ? (size_t)(r->nt - 1) : r->hkeep;
    size_t keep_from = r->pos > need ? r->pos - need : 0;
    if (keep_from > r->in.n) keep_from = r->in.n;          /* decimation larger than the arm: the next position lies in future input */
    if (keep_from > 0) { qv_drop(&r->in, keep_from); r->pos -= keep_from; }
}

/* fft_filter_ccc restated in direct form: complex taps, complex stream */
typedef struct { int nt; float* h; qvec in; size_t pos; size_t hkeep; } fircc_t;
static void fircc_init(fircc_t* f, const float* taps_c, int nt)
{
    f->nt = nt; f->h = (float*)malloc(sizeof(float) * 2 * nt);
    memcpy(f->h, taps_c, sizeof(float) * 2 * nt);
    qv_init(&f->in, 8); qv_push_zero(&f->in, nt - 1); f->pos = nt - 1; f->hkeep = 0;
}
static void fircc_retap(fircc_t* f, const float* taps_c, int nt)      /* see resamp_retap */
{
    free(f->h);
    f->nt = nt; f->h = (float*)malloc(sizeof(float) * 2 * nt);
    memcpy(f->h, taps_c, sizeof(float) * 2 * nt);
    if (f->pos < (size_t)(nt - 1)) {
        const size_t add = (size_t)(nt - 1) - f->pos, n0 = f->in.n;
        qv_push_zero(&f->in, add);
        memmove(f->in.d + add * 8, f->in.d, n0 * 8);
        memset(f->in.d, 0, add * 8);
        f->pos += add;
    }
}
static void fircc_free(fircc_t* f) { free(f->h); qv_free(&f->in); }
static void fircc_work(fircc_t* f, const float* x, size_t n, qvec* out)
{
    qv_push(&f->in, x, n);
    const float* b = (const float*)f->in.d;
    while (f->pos < f->in.n) {
        float re = 0.0f, im = 0.0f;
        for (int j = f->nt - 1; j >= 0; j--) {
            float hr = f->h[2 * j], hi = f->h[2 * j + 1];
            float xr = b[2 * (f->pos - j)], xi = b[2 * (f->pos - j) + 1];
            re = fmaf(hr, xr, re); re = fmaf(-hi, xi, re);
            im = fmaf(hr, xi, im); im = fmaf(hi, xr, im);
        }
        qv_pushc(out, re, im);
        f->pos++;
    }
    const size_t need = (size_t)(f->nt - 1) > f->hkeep ? (size_t)(f->nt - 1) : f->hkeep;
    size_t keep_from = f->pos > need ? f->pos - need : 0;
    if (keep_from > 0) { qv_drop(&f->in, keep_from); f->pos -= keep_from; }
}

/* ------------------------------------------------------------------ analog blocks */
/* analog::quadrature_demod_cf (A4) */
typedef struct { float gain; float pr, pi; } qdemod_t;
static void qdemod_init(qdemod_t* q, float gain) { q->gain = gain; q->pr = 0; q->pi = 0; }
static void qdemod_work(qdemod_t* q, const float* x, size_t n, qvec* out)
{
    for (size_t i = 0; i < n; i++) {
        float ar = x[2 * i], ai = x[2 * i + 1];
        float re = ar * q->pr + ai * q->pi;      /* x[n] * conj(x[n-1]) */
        float im = ai * q->pr - ar * q->pi;
        qv_pushf(out, q->gain * qo_fast_atan2f(im, re));
        q->pr = ar; q->pi = ai;
    }
}
/* analog::pwr_squelch_cc / squelch_base_cc (A5) */
typedef struct { double alpha, pwr, threshold; int ramp, ramped, state, gate; double envelope; } squelch_t;
enum { SQ_MUTED, SQ_ATTACK, SQ_UNMUTED, SQ_DECAY };
static void squelch_init(squelch_t* s, double db, double alpha, int ramp, int gate)
{
    s->alpha = alpha; s->pwr = 0; s->threshold = pow(10.0, db / 10.0); s->ramp = ramp; s->ramped = 0;
    s->state = SQ_MUTED; s->gate = gate; s->envelope = ramp ? 0.0 : 1.0;
}
static void squelch_work(squelch_t* s, const float* x, size_t n, qvec* out)
{
    for (size_t i = 0; i < n; i++) {
        float re = x[2 * i], im = x[2 * i + 1];
        float mag2 = re * re + im * im;
        s->pwr = s->alpha * (double)mag2 + (1.0 - s->alpha) * s->pwr;
        int mute = s->pwr < s->threshold;
        switch (s->state) {
        case SQ_MUTED: if (!mute) s->state = s->ramp ? SQ_ATTACK : SQ_UNMUTED; break;
        case SQ_UNMUTED: if (mute) s->state = s->ramp ? SQ_DECAY : SQ_MUTED; break;
        case SQ_ATTACK:
            s->envelope = 0.5 - cos(M_PI * (++s->ramped) / s->ramp) / 2.0;
            if (s->ramped >= s->ramp) { s->state = SQ_UNMUTED; s->envelope = 1.0; }
            break;
        case SQ_DECAY:
            s->envelope = 0.5 - cos(M_PI * (--s->ramped) / s->ramp) / 2.0;
            if (s->ramped == 0) s->state = SQ_MUTED;
            break;
        }
        if (s->state != SQ_MUTED) {
            float e = (float)s->envelope;
            qv_pushc(out, re * e, im * e);
        } else if (!s->gate) qv_pushc(out, 0.0f, 0.0f);
    }
}
/* analog::ctcss_squelch_ff (gr_demod_nbfm.cpp:60: make(8000, 88.5, 0.01, 8000, 160, true), switched in by set_ctcss(f != 0), :97-121).
 * GNU Radio's own block (gr-analog ctcss_squelch_ff_impl.cc + gr-fft goertzel.cc, not in /root/reference: restated, parity unpinned):
 * three Goertzel filters (float state) at the tone and its neighbours -- the adjacent entries of the standard CTCSS table, or
 * -/+2 % for a non-standard or edge tone -- over blocks of `len` items; at the end of a block the three magnitudes (rounded down to
 * 1e-5) decide: mute unless the tone's is at least `level` and not below either neighbour's; squelch_base_ff then gates / ramps. */
static const float ctcss_tones[38] = { 67.0f, 71.9f, 74.4f, 77.0f, 79.7f, 82.5f, 85.4f, 88.5f, 91.5f, 94.8f, 97.4f, 100.0f, 103.5f, 107.2f,
                                       110.9f, 114.8f, 118.8f, 123.0f, 127.3f, 131.8f, 136.5f, 141.3f, 146.2f, 151.4f, 156.7f, 162.2f, 167.9f,
                                       173.8f, 179.9f, 186.2f, 192.8f, 203.5f, 210.7f, 218.1f, 225.7f, 233.6f, 241.8f, 250.3f };
typedef struct { float wr, wi, d1, d2; int processed; } goertzel_t;
static void goertzel_init(goertzel_t* g, int rate, float freq)
{
    const float w = (float)(2.0 * M_PI * freq / rate);
    g->wr = (float)(2.0 * cosf(w)); g->wi = sinf(w); g->d1 = g->d2 = 0.0f; g->processed = 0;      /* std::cos / std::sin of a float */
}
static inline void goertzel_in(goertzel_t* g, float x)
{
    float y = x + g->wr * g->d1;
    y = y - g->d2;
    g->d2 = g->d1; g->d1 = y; g->processed++;
}
static inline float goertzel_mag(goertzel_t* g, int len)       /* |output()|, state reset */
{
    const float re = (float)((0.5 * g->wr * g->d1 - g->d2) / len), im = (g->wi * g->d1) / len;
    g->d1 = g->d2 = 0.0f; g->processed = 0;
    return (float)sqrt((double)re * re + (double)im * im);        /* std::abs(complex<float>) = hypotf */
}
typedef struct { int rate, len, ramp, ramped, state, gate, mute; float freq, level; double envelope; goertzel_t gl, gc, gr; } ctcss_t;
static void ctcss_set_frequency(ctcss_t* s, float freq)
{
    int idx = -1;
    for (int i = 0; i < 38; i++) if (ctcss_tones[i] == freq) idx = i;
    const float fl = (idx == -1 || idx == 0) ? (float)(freq * 0.98) : ctcss_tones[idx - 1];
    const float fr = (idx == -1 || idx == 37) ? (float)(freq * 1.02) : ctcss_tones[idx + 1];
    s->freq = freq;
    goertzel_init(&s->gl, s->rate, fl); goertzel_init(&s->gc, s->rate, freq); goertzel_init(&s->gr, s->rate, fr);
}
static void ctcss_init(ctcss_t* s, int rate, float freq, float level, int len, int ramp, int gate)
{
    memset(s, 0, sizeof *s);
    s->rate = rate; s->level = level; s->len = len; s->ramp = ramp; s->gate = gate; s->mute = 1;
    s->state = SQ_MUTED; s->envelope = ramp ? 0.0 : 1.0;
    ctcss_set_frequency(s, freq);
}
static void ctcss_work(ctcss_t* s, const float* x, size_t n, qvec* out)
{
    for (size_t i = 0; i < n; i++) {
        goertzel_in(&s->gl, x[i]); goertzel_in(&s->gc, x[i]); goertzel_in(&s->gr, x[i]);
        if (s->gc.processed == s->len) {
            const float rounder = 100000;
            float ml = goertzel_mag(&s->gl, s->len), mc = goertzel_mag(&s->gc, s->len), mr = goertzel_mag(&s->gr, s->len);
            ml = floorf(rounder * ml) / rounder; mc = floorf(rounder * mc) / rounder; mr = floorf(rounder * mr) / rounder;
            s->mute = (mc < s->level || mc < ml || mc < mr);
        }
        switch (s->state) {
        case SQ_MUTED: if (!s->mute) s->state = s->ramp ? SQ_ATTACK : SQ_UNMUTED; break;
        case SQ_UNMUTED: if (s->mute) s->state = s->ramp ? SQ_DECAY : SQ_MUTED; break;
        case SQ_ATTACK:
            s->envelope = 0.5 - cos(M_PI * (++s->ramped) / s->ramp) / 2.0;
            if (s->ramped >= s->ramp) { s->state = SQ_UNMUTED; s->envelope = 1.0; }
            break;
        case SQ_DECAY:
            s->envelope = 0.5 - cos(M_PI * (--s->ramped) / s->ramp) / 2.0;
            if (s->ramped == 0) s->state = SQ_MUTED;
            break;
        }
        if (s->state != SQ_MUTED) qv_pushf(out, (float)((double)x[i] * s->envelope));
        else if (!s->gate) qv_pushf(out, 0.0f);
    }
}
/* filter::iir_filter_ffd, 2-tap ff / 2-tap fb, oldstyle=false (A6) */
typedef struct { double b0, b1, a1; double x1, y1; } iir1_t;
static void iir1_init(iir1_t* f, const double* b, const double* a) { f->b0 = b[0]; f->b1 = b[1]; f->a1 = a[1]; f->x1 = 0; f->y1 = 0; }
static void iir1_work(iir1_t* f, const float* x, size_t n, qvec* out, float post_gain)
{
    for (size_t i = 0; i < n; i++) {
        double xin = (double)x[i];
        double acc = f->b0 * xin;
        acc = acc + f->b1 * f->x1;
        acc = acc - f->a1 * f->y1;
        f->x1 = xin; f->y1 = acc;
        qv_pushf(out, (float)acc * post_gain);
    }
}
/* analog::agc2_cc (A7).  GNU Radio's agc2.h is not in /root/reference (parity unpinned for this block): as restated here the COMPLEX
 * kernel picks the attack rate with a signed compare, `if ((tmp) > _gain)`, while the float kernel agc2_ff (AM chain below) uses
 * `fabsf(tmp) > _gain`.  The two forms only differ while gain < reference, i.e. for channel input levels above the reference, and
 * only for blocks whose attack and decay rates differ (QPSK: 1 / 0.1; SSB after set_agc_attack / set_agc_decay). */
typedef struct { float attack, decay, ref, gain, max_gain; } agc2_t;
static void agc2_init(agc2_t* a, float attack, float decay, float ref, float gain) { a->attack = attack; a->decay = decay; a->ref = ref; a->gain = gain; a->max_gain = 65536.0f; }
static inline void agc2_step(agc2_t* a, float xr, float xi, float* yr, float* yi)
{
    float orr = xr * a->gain, oi = xi * a->gain;
    float tmp = -a->ref + sqrtf(orr * orr + oi * oi);
    float rate = a->decay;
    if (tmp > a->gain) rate = a->attack;
    a->gain -= tmp * rate;
    if (a->gain < 0.0f) a->gain = 10e-5f;
    if (a->max_gain > 0.0f && a->gain > a->max_gain) a->gain = a->max_gain;
    *yr = orr; *yi = oi;
}
/* blocks::control_loop + digital::costas_loop_cc (A8) */
typedef struct { float phase, freq, alpha, beta, max_freq, min_freq; int order, use_snr; } costas_t;
static void costas_init(costas_t* c, float loop_bw, int order, int use_snr)
{
    tabs_init();
    c->phase = 0; c->freq = 0; c->max_freq = 1.0f; c->min_freq = -1.0f; c->order = order; c->use_snr = use_snr;
    qo_control_loop_gains(loop_bw, &c->alpha, &c->beta);
}
static inline void costas_step(costas_t* c, float xr, float xi, float* yr, float* yi)
{
    float sn, cs;
    qo_sincosf(-c->phase, &sn, &cs);
    float orr = xr * cs - xi * sn;
    float oi = xr * sn + xi * cs;
    float err;
    if (c->order == 2) {
        if (c->use_snr) { float snr = orr * orr + oi * oi; err = tanhf_lut(snr * orr) * oi; }
        else err = orr * oi;
    } else {
        if (c->use_snr) {
            float snr = orr * orr + oi * oi;
            err = tanhf_lut(snr * orr) * oi - tanhf_lut(snr * oi) * orr;
        } else {
            err = (orr > 0.0f ? 1.0f : -1.0f) * oi - (oi > 0.0f ? 1.0f : -1.0f) * orr;
        }
    }
    err = clipf(err, 1.0f);
    c->freq = c->freq + c->beta * err;
    c->phase = c->phase + c->freq + c->alpha * err;
    while ((double)c->phase > 2.0 * M_PI) c->phase = (float)((double)c->phase - 2.0 * M_PI);
    while ((double)c->phase < -2.0 * M_PI) c->phase = (float)((double)c->phase + 2.0 * M_PI);
    if (c->freq > c->max_freq) c->freq = c->max_freq;
    else if (c->freq < c->min_freq) c->freq = c->min_freq;
    *yr = orr; *yi = oi;
}

/* ------------------------------------------------------------------ symbol_sync_{ff,cc} (A9) */
enum { SL_RECT4 = 0, SL_DQPSK = 1, SL_BPSK = 2 };
typedef struct {
    int ncomp, slicer;
    int ted_plain;          /* 0 = TED_MOD_MUELLER_AND_MULLER (default), 1 = TED_MUELLER_AND_MULLER (gr_demod_dmr.cpp:66) */
    float sps, alpha, beta, max_period, min_period;
    float avg_period, inst_period, mu;
    float xr[3], xi[3], dr[3], di[3];
    size_t ii;
    int lookahead;
    qvec in;
} symsync_t;
static void symsync_init(symsync_t* s, int ncomp, float sps, float loop_bw, float damping, float ted_gain, float max_dev, int slicer)
{
    tabs_init();
    memset(s, 0, sizeof *s);
    s->ncomp = ncomp; s->slicer = slicer; s->sps = sps;
    qo_clock_loop_gains(loop_bw, damping, ted_gain, &s->alpha, &s->beta);
    s->max_period = sps + max_dev; s->min_period = sps - max_dev;
    s->avg_period = sps; s->inst_period = sps; s->mu = 0.0f; s->ii = 0;
    s->lookahead = 8 + (int)ceilf(s->max_period) + 1;
    qv_init(&s->in, sizeof(float) * ncomp);
}
static void symsync_free(symsync_t* s) { qv_free(&s->in); }
static inline void slice(int slicer, float re, float im, float* dr, float* di)
{
    if (slicer == SL_RECT4) {
        /* digital::constellation_rect(points {-1.5,-0.5,0.5,1.5}, real_sectors 4, width 1) */
        int sec = (int)floorf(re + 2.0f);
        if (sec < 0) sec = 0; if (sec > 3) sec = 3;
        *dr = -1.5f + (float)sec; *di = 0.0f;
    } else if (slicer == SL_DQPSK) {
        *dr = re > 0.0f ? 0.707107f : -0.707107f;
        *di = im > 0.0f ? 0.707107f : -0.707107f;
    } else {
        *dr = re > 0.0f ? 1.0f : -1.0f; *di = 0.0f;
    }
}
/* 8-tap MMSE interpolation, oldest sample first: sum_k taps[imu][7-i] * in[i] */
static inline float mmse_interp(const float* in, int stride, float mu)
{
    int imu = (int)rintf(mu * 128.0f);
    const float* t = g_mmse_tab + imu * 8;
    float acc = 0.0f;
    for (int i = 0; i < 8; i++) acc = fmaf(t[7 - i], in[i * stride], acc);
    return acc;
}
/* out: interpolated symbols (float or complex). */
static void symsync_work(symsync_t* s, const float* x, size_t n, qvec* out)
{
    qv_push(&s->in, x, n);
    const float* b = (const float*)s->in.d;
    while (s->ii + (size_t)s->lookahead <= s->in.n) {
        const float* p = b + s->ii * s->ncomp;
        float yr, yi = 0.0f;
        if (s->ncomp == 2) { yr = mmse_interp(p, 2, s->mu); yi = mmse_interp(p + 1, 2, s->mu); }
        else yr = mmse_interp(p, 1, s->mu);
        /* TED input */
        s->xr[2] = s->xr[1]; s->xr[1] = s->xr[0]; s->xr[0] = yr;
        s->xi[2] = s->xi[1]; s->xi[1] = s->xi[0]; s->xi[0] = yi;
        s->dr[2] = s->dr[1]; s->dr[1] = s->dr[0];
        s->di[2] = s->di[1]; s->di[1] = s->di[0];
        slice(s->slicer, yr, yi, &s->dr[0], &s->di[0]);
        float err;
        if (s->ncomp == 2) {
            /* TED_MOD_MUELLER_AND_MULLER complex: Re{(x0-x2) conj(d1) - (d0-d2) conj(x1)} */
            float ar = s->xr[0] - s->xr[2], ai = s->xi[0] - s->xi[2];
            float br = s->dr[0] - s->dr[2], bi = s->di[0] - s->di[2];
            float u = (ar * s->dr[1] + ai * s->di[1]) - (br * s->xr[1] + bi * s->xi[1]);
            err = clipf(u, 1.0f);
        } else if (s->ted_plain) {
            /* TED_MUELLER_AND_MULLER (real input): e = d[n-1] x[n] - d[n] x[n-1], clipped to +-1 (timing_error_detector_type.cc) */
            float u = s->dr[1] * s->xr[0] - s->dr[0] * s->xr[1];
            err = clipf(u, 1.0f);
        } else {
            float u = (s->xr[0] - s->xr[2]) * s->dr[1] - (s->dr[0] - s->dr[2]) * s->xr[1];
            err = clipf(u / 2.0f, 1.0f);
        }
        /* clock_tracking_loop::advance_loop */
        s->avg_period = s->avg_period + s->beta * err;
        if (s->avg_period > s->max_period) s->avg_period = s->max_period;
        else if (s->avg_period < s->min_period) s->avg_period = s->min_period;
        s->inst_period = s->avg_period + s->alpha * err;
        if (s->inst_period <= 0.0f) s->inst_period = s->avg_period;
        float ph = s->mu + s->inst_period;
        float fl = floorf(ph);
        s->mu = ph - fl;
        s->ii += (size_t)(int)fl;
        if (s->ncomp == 2) qv_pushc(out, yr, yi); else qv_pushf(out, yr);
    }
    /* drop consumed input */
    if (s->ii > 0) {
        size_t d = s->ii < s->in.n ? s->ii : s->in.n;
        qv_drop(&s->in, d); s->ii -= d;
    }
}

/* ------------------------------------------------------------------ fll_band_edge_cc (A8) */
/* digital::fll_band_edge_cc(sps, rolloff, filter_size, bw): rotate by the NCO, band-edge filter pair on the rotated
 * stream, error = |lower|^2 - |upper|^2 drives a second-order control loop (max/min freq = +-2*pi*2/sps).
 * Dot products: taps index k ascending = oldest rotated sample first, complex multiply-accumulate in four fmaf chains. */
typedef struct {
    int N; float* lo; float* up;           /* complex taps, interleaved, stored reversed like d_taps_lower/upper */
    float* hist;                            /* last N rotated outputs (complex), hist[N-1] newest */
    float phase, freq, alpha, beta, max_freq, min_freq;
} fll_t;
static double fll_sinc(double x) { if (x == 0) return 1.0; return sin(M_PI * x) / (M_PI * x); }
static void fll_init(fll_t* f, float sps, float rolloff, int N, float bw)
{
    f->N = N; f->lo = (float*)calloc(2 * N, 4); f->up = (float*)calloc(2 * N, 4); f->hist = (float*)calloc(2 * N, 4);
    f->phase = 0; f->freq = 0;
    qo_control_loop_gains(bw, &f->alpha, &f->beta);
    f->max_freq = (float)(2.0 * M_PI * (2.0 / sps)); f->min_freq = -f->max_freq;
    int M = (int)rint(N / sps);
    float power = 0;
    float* bb = (float*)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) {
        float k = (float)(-M + i * 2.0 / sps);
        float tap = (float)(fll_sinc(rolloff * k - 0.5) + fll_sinc(rolloff * k + 0.5));
        power += tap; bb[i] = tap;
    }
    int Nh = (int)((N - 1.0) / 2.0);
    for (int i = 0; i < N; i++) {
        float tap = bb[i] / power;
        float k = (float)((-Nh + i) / (2.0 * sps));
        float ang = (float)(2.0 * M_PI * (1 + rolloff) * k);
        f->lo[2 * (N - i - 1)] = tap * cosf(-ang); f->lo[2 * (N - i - 1) + 1] = tap * sinf(-ang);
        f->up[2 * (N - i - 1)] = tap * cosf(ang);  f->up[2 * (N - i - 1) + 1] = tap * sinf(ang);
    }
    free(bb);
}
static void fll_free(fll_t* f) { free(f->lo); free(f->up); free(f->hist); }
static inline void fll_step(fll_t* f, float xr, float xi, float* yr, float* yi)
{
    if (g_dbg_bypass_fll) { *yr = xr; *yi = xi; return; }
    float sn, cs;
    qo_sincosf(f->phase, &sn, &cs);
    float orr = xr * cs - xi * sn, oi = xr * sn + xi * cs;
    memmove(f->hist, f->hist + 2, sizeof(float) * 2 * (f->N - 1));
    f->hist[2 * (f->N - 1)] = orr; f->hist[2 * (f->N - 1) + 1] = oi;
    float ur = 0, ui = 0, lr = 0, li = 0;
    for (int k = 0; k < f->N; k++) {
        float hr = f->hist[2 * k], hi = f->hist[2 * k + 1];
        ur = fmaf(f->up[2 * k], hr, ur); ur = fmaf(-f->up[2 * k + 1], hi, ur);
        ui = fmaf(f->up[2 * k], hi, ui); ui = fmaf(f->up[2 * k + 1], hr, ui);
        lr = fmaf(f->lo[2 * k], hr, lr); lr = fmaf(-f->lo[2 * k + 1], hi, lr);
        li = fmaf(f->lo[2 * k], hi, li); li = fmaf(f->lo[2 * k + 1], hr, li);
    }
    float err = (lr * lr + li * li) - (ur * ur + ui * ui);
    f->freq = f->freq + f->beta * err;
    f->phase = f->phase + f->freq + f->alpha * err;
    while ((double)f->phase > 2.0 * M_PI) f->phase = (float)((double)f->phase - 2.0 * M_PI);
    while ((double)f->phase < -2.0 * M_PI) f->phase = (float)((double)f->phase + 2.0 * M_PI);
    if (f->freq > f->max_freq) f->freq = f->max_freq;
    else if (f->freq < f->min_freq) f->freq = f->min_freq;
    *yr = orr; *yi = oi;
}

/* ------------------------------------------------------------------ clock_recovery_mm_cc (A9) */
typedef struct {
    float mu, omega, omega_mid, omega_lim, gain_omega, gain_mu;
    float p0r, p0i, p1r, p1i, p2r, p2i, c0r, c0i, c1r, c1i, c2r, c2i;
    size_t ii; qvec in;
} crmm_t;
static void crmm_init(crmm_t* c, float omega, float gain_omega, float mu, float gain_mu, float rel_lim)
{
    tabs_init();
    memset(c, 0, sizeof *c);
    c->mu = mu; c->omega = omega; c->omega_mid = omega; c->omega_lim = rel_lim * omega;
    c->gain_omega = gain_omega; c->gain_mu = gain_mu;
    qv_init(&c->in, 8);
}
static void crmm_work(crmm_t* c, const float* x, size_t n, qvec* out)
{
    qv_push(&c->in, x, n);
    const float* b = (const float*)c->in.d;
    while (c->ii + 24 <= c->in.n) {                     /* 8 interpolator taps + FUDGE 16 */
        const float* p = b + 2 * c->ii;
        c->p2r = c->p1r; c->p2i = c->p1i; c->p1r = c->p0r; c->p1i = c->p0i;
        c->p0r = mmse_interp(p, 2, c->mu); c->p0i = mmse_interp(p + 1, 2, c->mu);
        c->c2r = c->c1r; c->c2i = c->c1i; c->c1r = c->c0r; c->c1i = c->c0i;
        c->c0r = c->p0r > 0.0f ? 1.0f : 0.0f; c->c0i = c->p0i > 0.0f ? 1.0f : 0.0f;   /* slicer_0deg */
        /* x = (c0 - c2) * conj(p1);  y = (p0 - p2) * conj(c1);  mm = Re(y - x) */
        float ar = c->c0r - c->c2r, ai = c->c0i - c->c2i;
        float xr = ar * c->p1r + ai * c->p1i;
        float br = c->p0r - c->p2r, bi = c->p0i - c->p2i;
        float yr = br * c->c1r + bi * c->c1i;
        float mm = clipf(yr - xr, 1.0f);
        qv_pushc(out, c->p0r, c->p0i);
        c->omega = c->omega + c->gain_omega * mm;
        c->omega = c->omega_mid + clipf(c->omega - c->omega_mid, c->omega_lim);
        c->mu = c->mu + c->omega + c->gain_mu * mm;
        float fl = floorf(c->mu);
        c->ii += (size_t)(int)fl;
        c->mu = c->mu - fl;
    }
    if (c->ii > 0) { size_t d = c->ii < c->in.n ? c->ii : c->in.n; qv_drop(&c->in, d); c->ii -= d; }
}

/* ------------------------------------------------------------------ FEC (A11) and LFSR (A12) */
static inline int parity8(unsigned v) { v ^= v >> 4; v ^= v >> 2; v ^= v >> 1; return v & 1; }
typedef struct { unsigned state; } ccenc_t;
static void ccenc_work(ccenc_t* e, const uint8_t* bits, size_t n, qvec* out)
{
    for (size_t i = 0; i < n; i++) {
        e->state = ((e->state << 1) | (bits[i] & 1)) & 0x7f;
        qv_pushb(out, (uint8_t)parity8(e->state & 109));
        qv_pushb(out, (uint8_t)parity8(e->state & 79));
    }
}
/* fec::decoder(cc_decoder(80,7,2,{109,79},0,-1,CC_STREAMING)) : 8-bit metrics, generic VOLK butterfly */
typedef struct { qvec in; int start_state; unsigned char branchtab[64]; } ccdec_t;
static void ccdec_init(ccdec_t* d)
{
    qv_init(&d->in, 1);
    qv_push_zero(&d->in, 12); /* history = rate*(k-1) zero items */
    d->start_state = 0;
    int polys[2] = { 109, 79 };
    for (int st = 0; st < 32; st++)
        for (int i = 0; i < 2; i++) d->branchtab[i * 32 + st] = parity8((2 * st) & polys[i]) ? 255 : 0;
}
static void ccdec_free(ccdec_t* d) { qv_free(&d->in); }
static void ccdec_block(ccdec_t* d, const unsigned char* syms, unsigned char* out80)
{
    unsigned char m1[64], m2[64];
    unsigned char* X = m1; unsigned char* Y = m2;
    unsigned int dec[86][2];
    memset(dec, 0, sizeof dec);
    for (int i = 0; i < 64; i++) X[i] = 63;
    X[d->start_state & 63] = 0;
    for (int s = 0; s < 86; s++) {
        for (int i = 0; i < 32; i++) {
            unsigned char metric = 0, m0, mm1, mm2, m3;
            for (int j = 0; j < 2; j++) metric += (d->branchtab[i + j * 32] ^ syms[s * 2 + j]) >> 2;
            metric = metric >> 2;
            const unsigned char max = ((2 * ((256 - 1) >> 2)) >> 2);
            m0 = X[i] + metric;
            mm1 = X[i + 32] + (max - metric);
            mm2 = X[i] + (max - metric);
            m3 = X[i + 32] + metric;
            int decision0 = (signed int)(m0 - mm1) > 0;
            int decision1 = (signed int)(mm2 - m3) > 0;
            Y[2 * i] = decision0 ? mm1 : m0;
            Y[2 * i + 1] = decision1 ? m3 : mm2;
            dec[s][i / 16] |= (unsigned)(decision0 | decision1 << 1) << ((2 * i) & 31);
        }
        unsigned char min = Y[0];
        for (int i = 0; i < 64; i++) if (min > Y[i]) min = Y[i];
        for (int i = 0; i < 64; i++) Y[i] -= min;
        unsigned char* t = X; X = Y; Y = t;
    }
    /* find_endstate: minimum final metric, first index on ties */
    int endstate = 0; unsigned char best = X[0];
    for (int i = 1; i < 64; i++) if (X[i] < best) { best = X[i]; endstate = i; }
    /* chainback_viterbi(out, 80, endstate, tailsize 6) */
    unsigned es = (unsigned)(endstate % 64) << 2;
    int retval = 0;
    int nbits = 80;
    while (nbits-- > 80 - 6) {
        unsigned st = es >> 2;
        int k = (dec[nbits + 6][st / 32] >> (st % 32)) & 1;
        es = (es >> 1) | ((unsigned)k << (7 - 2 + 2));
        out80[nbits % 80] = (unsigned char)k;
        retval = (int)es;
    }
    nbits += 1;
    while (nbits-- != 0) {
        unsigned st = es >> 2;
        int k = (dec[nbits + 6][st / 32] >> (st % 32)) & 1;
        es = (es >> 1) | ((unsigned)k << (7 - 2 + 2));
        out80[nbits % 80] = (unsigned char)k;
    }
    d->start_state = retval >> 2;
}
static void ccdec_work(ccdec_t* d, const uint8_t* soft, size_t n, qvec* out)
{
    qv_push(&d->in, soft, n);
    size_t off = 0;
    while (d->in.n - off >= 172) {
        unsigned char o[80];
        ccdec_block(d, d->in.d + off, o);
        qv_push(out, o, 80);
        off += 160;
    }
    qv_drop(&d->in, off);
}
/* digital::lfsr(0x8A, 0x7F, 7) */
typedef struct { unsigned reg; } lfsr_t;
static void lfsr_init(lfsr_t* l) { l->reg = 0x7F; }
static inline unsigned char lfsr_scramble(lfsr_t* l, unsigned char in)
{
    unsigned char out = l->reg & 1;
    unsigned newbit = (parity8(l->reg & 0x8A) ^ (in & 1)) & 1;
    l->reg = ((l->reg >> 1) | (newbit << 7)) & 0xff;
    return out;
}
static inline unsigned char lfsr_descramble(lfsr_t* l, unsigned char in)
{
    unsigned char out = (unsigned char)(parity8(l->reg & 0x8A) ^ (in & 1));
    l->reg = ((l->reg >> 1) | ((unsigned)(in & 1) << 7)) & 0xff;
    return out;
}
/* blocks::float_to_uchar after multiply_const / add_const (A10) */
static inline unsigned char soft_u8(float v, float scale)
{
    float t = v * scale;
    t = t + 128.0f;
    float r = rintf(t);
    if (r < 0.0f) r = 0.0f; else if (r > 255.0f) r = 255.0f;
    return (unsigned char)r;
}

/* ------------------------------------------------------------------ stand-alone helpers */
long qo_fir_decim_ccf(const float* h, int ntaps, int D, const float* x, long n, float* y)
{
    resamp_t r; qvec out; qv_init(&out, 8);
    resamp_init(&r, 2, 1, D, h, ntaps);
    resamp_work(&r, x, (size_t)n, &out);
    memcpy(y, out.d, out.n * 8);
    long k = (long)out.n;
    resamp_free(&r); qv_free(&out);
    return k;
}
long qo_fir_fff(const float* h, int ntaps, int L, int M, const float* x, long n, float* y, long ycap)
{
    resamp_t r; qvec out; qv_init(&out, 4);
    resamp_init(&r, 1, L, M, h, ntaps);
    resamp_work(&r, x, (size_t)n, &out);
    long k = (long)out.n; if (k > ycap) k = ycap;
    memcpy(y, out.d, (size_t)k * 4);
    resamp_free(&r); qv_free(&out);
    return k;
}
long qo_cc_encode(const uint8_t* bits, long n, uint8_t* outb)
{
    ccenc_t e = { 0 }; qvec out; qv_init(&out, 1);
    ccenc_work(&e, bits, (size_t)n, &out);
    memcpy(outb, out.d, out.n); long k = (long)out.n; qv_free(&out); return k;
}
long qo_cc_decode(const uint8_t* soft, long n, uint8_t* outb)
{
    ccdec_t d; ccdec_init(&d); qvec out; qv_init(&out, 1);
    ccdec_work(&d, soft, (size_t)n, &out);
    memcpy(outb, out.d, out.n); long k = (long)out.n; qv_free(&out); ccdec_free(&d); return k;
}
void qo_scramble(const uint8_t* in, long n, uint8_t* out) { lfsr_t l; lfsr_init(&l); for (long i = 0; i < n; i++) out[i] = lfsr_scramble(&l, in[i]); }
void qo_descramble(const uint8_t* in, long n, uint8_t* out) { lfsr_t l; lfsr_init(&l); for (long i = 0; i < n; i++) out[i] = lfsr_descramble(&l, in[i]); }

/* ------------------------------------------------------------------ RX chains */
/* gr::dsss::dsss_decoder_cc (/root/reference/src/gr/dsss_decoder_cc_impl.cc:45-175): matched filter = the reversed code, `samples`
 * items per chip, RRC-shaped (:60-96: N + rrc_ntaps taps, rrc_ntaps = 11 * samples); per output symbol m the N = samples * code_len
 * correlations y_j = fir_filter_ccc(taps).filter(in + (i - 1) N + j), j = 0..N-1, and the strongest one (first strict maximum of
 * |y|) times 2 / N goes out (:150-166).  With history N the pointer arithmetic reads from item m N - 2 N + 1 + j on: N items
 * BEFORE the declared history.  In GNU Radio that region is whatever the circular buffer still holds; HERE it is defined as the
 * stream's own older items (zeros before the stream began), which makes the block a chunk-invariant stream function.
 * |y| = std::abs(complex<float>) = hypotf, glibc: (float)sqrt((double)re * re + (double)im * im). */
typedef struct { int N, ntaps; float* tr; qvec in; long long n_in, m_out; } dsssdec_t;
void qo_dsss_decoder_taps(const int* code, int code_len, int samples, float* taps_c /* [(N + 11 samples)][2] */)
{
    const int N = samples * code_len, extra = samples * 11, total = N + 2 * extra;
    float* cs = (float*)calloc((size_t)total, sizeof(float));
    for (int i = 0; i < code_len; i++) {
        const float c = code[code_len - (i + 1)] == 0 ? -1.0f : 1.0f;
        for (int k = 0; k < samples; k++) cs[extra + i * samples + k] = c;
    }
    static float rrc[16384];
    const int nr = qo_firdes_rrc(1, samples, 1.0, 0.350f, extra, rrc, 16384);
    /* fir_filter_ccf(rrc).filter(&code_symbols[i]) = sum_k rrc[nr-1-k] * cs[i+k]; code symbols are real: imaginary part 0 */
    for (int i = 0; i < N + extra; i++) {
        float acc = 0.0f, acci = 0.0f;
        for (int k = 0; k < nr; k++) { acc = acc + cs[i + k] * rrc[nr - 1 - k]; acci = acci + 0.0f * rrc[nr - 1 - k]; }
        taps_c[2 * i] = acc; taps_c[2 * i + 1] = acci;
    }
    free(cs);
}
static void dsssdec_init(dsssdec_t* d, const int* code, int code_len, int samples)
{
    memset(d, 0, sizeof *d);
    d->N = samples * code_len; d->ntaps = d->N + samples * 11;
    float* t = (float*)malloc(sizeof(float) * 2 * (size_t)d->ntaps);
    qo_dsss_decoder_taps(code, code_len, samples, t);
    d->tr = (float*)malloc(sizeof(float) * 2 * (size_t)d->ntaps);           /* fir_filter_ccc stores the taps reversed */
    for (int k = 0; k < d->ntaps; k++) { d->tr[2 * k] = t[2 * (d->ntaps - 1 - k)]; d->tr[2 * k + 1] = t[2 * (d->ntaps - 1 - k) + 1]; }
    free(t);
    qv_init(&d->in, 8);
    /* the stream begins with 2 N - 1 zeros in front: in.d[i] is stream item i - (2 N - 1) + (items dropped so far) */
    float z[2] = { 0.0f, 0.0f };
    for (int i = 0; i < 2 * d->N - 1; i++) qv_push(&d->in, z, 1);
}
static void dsssdec_work(dsssdec_t* d, const float* x, size_t n, qvec* out)
{
    qv_push(&d->in, x, n);
    d->n_in += (long long)n;
    const int N = d->N;
    /* output m reads stream items up to m N + (ntaps - N) - 1: produced once they have arrived */
    while ((d->m_out * N + (d->ntaps - N)) <= d->n_in) {
        const float* b = (const float*)d->in.d;        /* b[0] = stream item m_out N - (2 N - 1) */
        float best = 0.0f, br = 0.0f, bi = 0.0f;
        for (int j = 0; j < N; j++) {
            const float* w = b + 2 * (size_t)j;
            float ar = 0.0f, ai = 0.0f;
            for (int k = 0; k < d->ntaps; k++) {
                const float xr = w[2 * k], xi = w[2 * k + 1], tr = d->tr[2 * k], ti = d->tr[2 * k + 1];
                const float pr = xr * tr - xi * ti, pi = xr * ti + xi * tr;
                ar = ar + pr; ai = ai + pi;
            }
            const float mag = (float)sqrt((double)ar * (double)ar + (double)ai * (double)ai);
            if (mag > best) { best = mag; br = ar; bi = ai; }
        }
        const float sc = 2.0f / (float)N;
        qv_pushc(out, br * sc, bi * sc);
        /* drop N items: the next symbol's window starts N later */
        memmove(d->in.d, (char*)d->in.d + 8 * (size_t)N, 8 * (d->in.n - (size_t)N));
        d->in.n -= (size_t)N;
        d->m_out++;
    }
}
/* the decoder alone, fed in `chunk`-item pieces (tests: against the compiled reference block, and chunk invariance) */
long qo_dsss_decoder_run(const int* code, int code_len, int samples, const float* in_c, long n, long chunk, float* out_c, long cap)
{
    dsssdec_t d; dsssdec_init(&d, code, code_len, samples);
    qvec out; qv_init(&out, 8);
    for (long lo = 0; lo < n; lo += chunk) dsssdec_work(&d, in_c + 2 * lo, (size_t)((n - lo) < chunk ? (n - lo) : chunk), &out);
    const long m = (long)out.n < cap ? (long)out.n : cap;
    memcpy(out_c, out.d, 8 * (size_t)m);
    qv_free(&out); qv_free(&d.in); free(d.tr);
    return m;
}

struct qo_rx {
    int gmsk;               /* 2FSK branch running as gr_demod_gmsk */
    int m17;                /* 4FSK (fm) branch running as gr_demod_m17 */
    int kind, fm;
    int sym_sps, tsr;
    /* stages (not all used by every kind) */
    resamp_t resamp;        /* stage-1 rational resampler (ccf) */
    resamp_t filt;          /* fft_filter_ccf restated direct-form */
    qdemod_t qd;
    resamp_t shaping;       /* RRC (fff or ccf) */
    symsync_t ss;
    float pm_sens;
    float soft_scale;
    ccdec_t dec; lfsr_t descr;
    /* nbfm */
    squelch_t sq; resamp_t audio_rs; resamp_t audio_filt; iir1_t deemph;
    /* qpsk */
    agc2_t agc; costas_t pll, costas; float dp_r, dp_i; float rot_r, rot_i;
    /* 4fsk non-fm */
    fircc_t bp[4]; resamp_t symfilt;
    /* ssb */
    fircc_t ssb_bpf; resamp_t ssb_audio; float env_m2, env_m1; qvec s_clip; size_t st_pos;
    float if_gain;          /* gr_demod_ssb's multiply_const_cc(0.9) (set_gain) */
    int filter_width, flag;
    /* bpsk / 2fsk */
    fll_t fll; crmm_t crmm; ccdec_t dec2; lfsr_t descr2; int dec2_started;
    /* nbfm tone squelch */
    ctcss_t ctcss; int ctcss_on;
    /* dsss */
    resamp_t resamp_if; dsssdec_t dsss; qvec s_dsss;
    /* front-end rotator (gr_demod_base.cpp:57,180,1220-1225): Q32 NCO, phase = base + inc * (n - n_base) */
    uint32_t rot_inc, rot_base; long long rot_nbase, rot_n; qvec s_rot;
    /* scratch + ports */
    qvec s_res, s_filt, s_dem, s_rrc, s_sym, s_soft, s_bits, s_tmp, s_tmp2, s_bp[4];
    qvec port[4];
    float taps_store[4][4096]; int ntaps_store[4];
};

static void rx_common_init(qo_rx* r)
{
    qv_init(&r->s_res, 8); qv_init(&r->s_filt, 8); qv_init(&r->s_dem, 4); qv_init(&r->s_rrc, 4);
    qv_init(&r->s_sym, 4); qv_init(&r->s_soft, 1); qv_init(&r->s_bits, 1); qv_init(&r->s_tmp, 8); qv_init(&r->s_tmp2, 8);
    for (int i = 0; i < 4; i++) qv_init(&r->s_bp[i], 8);
    qv_init(&r->port[0], 8); qv_init(&r->port[1], 8); qv_init(&r->port[2], 1); qv_init(&r->port[3], 1);
}

qo_rx* qo_rx_create(int kind, int sps, int samp_rate, int carrier_freq, int filter_width, int flag)
{
    (void)carrier_freq;
    tabs_init();
    qo_rx* r = (qo_rx*)calloc(1, sizeof *r);
    r->kind = kind;
    r->if_gain = 0.9f; r->filter_width = filter_width; r->flag = flag;
    rx_common_init(r);
    float* T0 = r->taps_store[0]; float* T1 = r->taps_store[1]; float* T2 = r->taps_store[2]; float* T3 = r->taps_store[3];
    /* gr_demod_gmsk.cpp:30-134 is the 2FSK (fm) chain without the band-edge FLL, with a plain low-pass as symbol filter and
     * its own clock-loop constants: restated through the 2FSK branch */
    const int gmsk = (kind == QO_DEMOD_GMSK);
    if (gmsk) { kind = QO_DEMOD_2FSK; r->kind = QO_DEMOD_2FSK; r->gmsk = 1; flag = 1; }
    if (kind == QO_DEMOD_DMR) {
        /* /root/reference/src/gr/gr_demod_dmr.cpp:30-112 (oracle only so far: the CUDA path is not built): rational_resampler_ccf(3, 125)
         * with low_pass_2(3, 3 fs, 5000, 2000, 60, BH) -> [port 0 at 24 ksps] -> quadrature demod (24000 / (pi/2 * 4800)) ->
         * RRC(1, 24k, 4800, 0.2, 125) -> [port 3, float] -> symbol_sync_ff(TED_MUELLER_AND_MULLER, 5, 2 pi / 100, 1, 0.2869, 0.06, rect4)
         * -> x0.9 -> phase_modulator_fc(pi/2) -> [port 1]; re / im -> slicer -> pack 2 -> map {3,1,2,0} -> unpack 2 -> [port 2] */
        r->kind = QO_DEMOD_DMR; r->fm = 1; r->m17 = 1;
        r->tsr = 24000; r->sym_sps = 5;
        int n0 = qo_firdes_low_pass_2(3, 3.0 * samp_rate, 5000, 2000, 60, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 3, 125, T0, n0);
        const float symbol_rate = (float)r->tsr / (float)r->sym_sps;
        qdemod_init(&r->qd, (float)(r->tsr / (M_PI / 2 * symbol_rate)));
        int n2 = qo_firdes_rrc(1, r->tsr, r->tsr / r->sym_sps, 0.2, 25 * r->sym_sps, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->shaping, 1, 1, 1, T2, n2);
        symsync_init(&r->ss, 1, (float)r->sym_sps, (float)(2 * M_PI / 100.0f), 1.0f, 0.2869f, 0.06f, SL_RECT4);
        r->ss.ted_plain = 1;
        r->pm_sens = (float)(M_PI / 2);
        r->port[3].isz = 4;
    } else if (kind == QO_DEMOD_M17) {
        /* /root/reference/src/gr/gr_demod_m17.cpp:30-113: rational_resampler_ccf(3, 125) to 24 ksps -> low-pass -> quadrature
         * demod (5 / pi) -> RRC(1.5, 24k, 4800, 0.5, 250) -> symbol_sync_ff (4-level) -> phase_modulator_fc(pi/2) -> [port 1];
         * re / im -> binary_slicer -> pack 2 -> map {3,1,2,0} -> unpack 2 -> [port 2] (no FEC in this block).  Runs through the
         * 4FSK (fm) branch with its own constants and bit tail. */
        r->kind = QO_DEMOD_4FSK; r->fm = 1; r->m17 = 1;
        r->tsr = 24000; r->sym_sps = 5;
        int n0 = qo_firdes_low_pass(3, 3.0 * samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 3, 125, T0, n0);
        int n1 = qo_firdes_low_pass(1, r->tsr, filter_width, filter_width, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->filt, 2, 1, 1, T1, n1);
        qdemod_init(&r->qd, (float)(r->sym_sps / M_PI));
        int n2 = qo_firdes_rrc(1.5, r->tsr, r->tsr / r->sym_sps, 0.5, 50 * r->sym_sps, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->shaping, 1, 1, 1, T2, n2);
        const float symbol_rate = (float)r->tsr / (float)r->sym_sps;
        symsync_init(&r->ss, 1, (float)r->sym_sps, (float)(2 * M_PI / (symbol_rate / 50)), 1.0f, 0.2869f, 500.0f / symbol_rate, SL_RECT4);
        r->pm_sens = (float)(M_PI / 2);
        r->soft_scale = 128.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr);
    } else if (kind == QO_DEMOD_4FSK) {
        /* /root/reference/src/gr/gr_demod_4fsk.cpp:32-205 */
        int fm = flag; r->fm = fm;
        int rs = 0, bw = 0, decimation = 1, interpolation = 1, nfilts = 0;
        if (sps == 1) { r->tsr = 80000; r->sym_sps = sps * 8; decimation = 25; interpolation = 2; rs = 10000; bw = 4000; nfilts = 32 * r->sym_sps; }
        if (sps == 5) { r->tsr = 20000; r->sym_sps = sps * 2; decimation = 50; interpolation = 1; rs = 2000; bw = 4000; nfilts = 25 * r->sym_sps; }
        if (sps == 10) { r->tsr = 10000; r->sym_sps = sps; decimation = 100; interpolation = 1; rs = 1000; bw = 2000; nfilts = 25 * r->sym_sps; }
        if (sps == 2) { interpolation = 1; decimation = 2; r->sym_sps = 5; r->tsr = 500000; nfilts = 50 * r->sym_sps; }
        if ((nfilts % 2) == 0) nfilts += 1;
        int n0 = qo_firdes_low_pass(interpolation, (double)interpolation * samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, interpolation, decimation, T0, n0);
        int n1 = qo_firdes_low_pass(1, r->tsr, filter_width, filter_width / 2, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->filt, 2, 1, 1, T1, n1);
        if (fm) {
            qdemod_init(&r->qd, (float)(r->sym_sps / (1 * M_PI)));
            int n2 = qo_firdes_rrc(1.5, r->tsr, r->tsr / r->sym_sps, 0.2, nfilts, T2, 4096);
            r->ntaps_store[2] = n2;
            resamp_init(&r->shaping, 1, 1, 1, T2, n2);
            symsync_init(&r->ss, 1, (float)r->sym_sps, (float)(2 * M_PI / 200.0f), 1.0f, 0.2869f, 0.05f, SL_RECT4);
        } else {
            float tc[2 * 4096];
            int fw = filter_width;
            double lo[4] = { -fw, -fw + rs, 0, fw - rs }, hi[4] = { -fw + rs, 0, fw - rs, fw };
            for (int i = 0; i < 4; i++) {
                int n = qo_firdes_complex_band_pass(1, r->tsr, lo[i], hi[i], bw, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
                fircc_init(&r->bp[i], tc, n);
            }
            int n3 = qo_firdes_low_pass(1.0, r->tsr, r->tsr / r->sym_sps, r->tsr / r->sym_sps / 20, QO_WIN_BLACKMAN_HARRIS, T3, 4096);
            r->ntaps_store[3] = n3;
            resamp_init(&r->symfilt, 2, 1, 1, T3, n3);
            symsync_init(&r->ss, 2, (float)r->sym_sps, (float)(2 * M_PI / 200.0f), 1.0f, 0.2869f, 0.05f, SL_RECT4);
        }
        r->pm_sens = (float)(M_PI / 2);
        r->soft_scale = 128.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr);
    } else if (kind == QO_DEMOD_QPSK) {
        /* /root/reference/src/gr/gr_demod_qpsk.cpp:33-159 */
        int decimation, interpolation = 1; float costas_bw = (float)(M_PI / 200);
        if (sps > 4 && sps < 125) { decimation = 25; r->sym_sps = sps * 4 / 25; r->tsr = 40000; }
        else if (sps >= 125) { decimation = 100; r->sym_sps = sps / 25; r->tsr = 10000; }
        else { decimation = 2; r->sym_sps = sps; r->tsr = 500000; costas_bw = (float)(M_PI / 400); }
        int n0 = qo_firdes_low_pass_2(interpolation, (double)samp_rate * interpolation, r->tsr / 2, r->tsr / 10, 60, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, interpolation, decimation, T0, n0);
        int n1 = qo_firdes_rrc(r->sym_sps, r->sym_sps, 1, 0.35, 11 * r->sym_sps, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->shaping, 2, 1, 1, T1, n1);
        agc2_init(&r->agc, 1.0f, 1e-1f, 1.0f, 1.0f);
        float symbol_rate = (float)r->tsr / (float)r->sym_sps;
        float sps_dev = 200.0f / symbol_rate;
        symsync_init(&r->ss, 2, (float)r->sym_sps, (float)(2 * M_PI / (symbol_rate / 10)), 1.0f, 0.2869f, sps_dev, SL_DQPSK);
        costas_init(&r->pll, (float)(M_PI / 200 / r->sym_sps), 4, 1);
        costas_init(&r->costas, costas_bw, 4, 1);
        r->dp_r = 0; r->dp_i = 0;
        float th = (float)(-3 * M_PI / 4);
        r->rot_r = cosf(th); r->rot_i = sinf(th);
        r->soft_scale = 48.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr);
        r->fm = (sps > 4); /* gr_demod_qpsk.cpp:130-138: the FLL is only connected for sps > 4 */
        if (r->fm) fll_init(&r->fll, (float)r->sym_sps, 0.35f, 32, (float)(2 * M_PI / 100));
    } else if (kind == QO_DEMOD_NBFM) {
        /* /root/reference/src/gr/gr_demod_nbfm.cpp:31-79 */
        r->tsr = 20000;
        double a[2], b[2];
        qo_deemph_taps(r->tsr, 50e-6, a, b);
        iir1_init(&r->deemph, b, a);
        int n0 = qo_firdes_low_pass(1, samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, 50, T0, n0);
        int n1 = qo_firdes_low_pass_2(1, r->tsr, filter_width, 3500, 60, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->filt, 2, 1, 1, T1, n1);
        int n2 = qo_firdes_low_pass_2(2, 2 * r->tsr, 3600, 250, 60, QO_WIN_BLACKMAN_HARRIS, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->audio_rs, 1, 2, 5, T2, n2);
        int n3 = qo_firdes_low_pass_2(1, 8000, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, T3, 4096);
        r->ntaps_store[3] = n3;
        resamp_init(&r->audio_filt, 1, 1, 1, T3, n3);
        qdemod_init(&r->qd, (float)(r->tsr / (4 * M_PI * filter_width)));
        squelch_init(&r->sq, -140, 0.01, 320, 1);
        ctcss_init(&r->ctcss, 8000, 88.5f, 0.01f, 8000, 160, 1);                    /* gr_demod_nbfm.cpp:60 */
        r->port[1].isz = 4;
    } else if (kind == QO_DEMOD_WBFM) {
        /* /root/reference/src/gr/gr_demod_wbfm.cpp:28-70: /5 (low_pass(1, fs, 100k, 100k, BH)) -> low_pass_2(1, 200k, fw, 600, 90, BH)
         * -> [port 0 at 200 ksps] -> pwr_squelch_cc(-140, 0.01, 0, gate) -> quadrature_demod_cf(200k / (2 pi fw)) -> x0.9 ->
         * iir_filter_ffd(de-emphasis taps designed for fs = 8000: the reference's own choice, :39-41) -> rational_resampler_fff(1, 25)
         * with low_pass(1, 200k, 4000, 2000, BH) -> [port 1 at 8 ksps] */
        r->tsr = 200000;
        double a[2], b[2];
        qo_deemph_taps(8000, 50e-6, a, b);
        iir1_init(&r->deemph, b, a);
        int n0 = qo_firdes_low_pass(1, samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, 5, T0, n0);
        int n1 = qo_firdes_low_pass_2(1, r->tsr, filter_width, 600, 90, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->filt, 2, 1, 1, T1, n1);
        int n2 = qo_firdes_low_pass(1, r->tsr, 4000, 2000, QO_WIN_BLACKMAN_HARRIS, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->audio_rs, 1, 1, 25, T2, n2);
        qdemod_init(&r->qd, (float)(r->tsr / (2 * M_PI * filter_width)));
        squelch_init(&r->sq, -140, 0.01, 0, 1);
        r->port[1].isz = 4;
    } else if (kind == QO_DEMOD_AM) {
        /* /root/reference/src/gr/gr_demod_am.cpp:28-82: /50 (419 taps) -> complex band-pass(-fw, fw) -> [port 0] ->
         * pwr_squelch_cc(-140, 0.01, 0, gate) -> complex_to_mag -> agc2_ff(.1, .1, 1, 1) -> iir_filter_ffd({1,-1},{0,.9999})
         * (old style: y = x - x1 + 0.9999 y1, double) -> x0.99 -> rational_resampler_fff(2,5) -> audio low-pass -> [port 1] */
        r->tsr = 20000;
        int n0 = qo_firdes_low_pass(1, samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, 50, T0, n0);
        float tc[2 * 4096];
        int nb = qo_firdes_complex_band_pass_2(1, r->tsr, -filter_width, filter_width, 200, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
        fircc_init(&r->ssb_bpf, tc, nb);
        squelch_init(&r->sq, -140, 0.01, 0, 1);
        agc2_init(&r->agc, 1e-1f, 1e-1f, 1.0f, 1.0f);
        { const double b[2] = { 1.0, -1.0 }, a[2] = { 1.0, -0.9999 }; iir1_init(&r->deemph, b, a); }
        int n2 = qo_firdes_low_pass(2, 2 * r->tsr, 3600, 600, QO_WIN_BLACKMAN_HARRIS, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->audio_rs, 1, 2, 5, T2, n2);
        int n3 = qo_firdes_low_pass(1, 8000, 3600, 300, QO_WIN_BLACKMAN_HARRIS, T3, 4096);
        r->ntaps_store[3] = n3;
        resamp_init(&r->audio_filt, 1, 1, 1, T3, n3);
        r->port[1].isz = 4;
    } else if (kind == QO_DEMOD_SSB) {
        /* /root/reference/src/gr/gr_demod_ssb.cpp:31-86; flag = sb (0 = USB, 1 = LSB) */
        r->tsr = 8000;
        int n0 = qo_firdes_low_pass(1, samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, sps, T0, n0);
        float tc[2 * 4096];
        int nb = flag ? qo_firdes_complex_band_pass_2(1, r->tsr, -filter_width, -200, 200, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096)
                      : qo_firdes_complex_band_pass_2(1, r->tsr, 200, filter_width, 200, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
        fircc_init(&r->ssb_bpf, tc, nb);
        squelch_init(&r->sq, -140, 0.01, 0, 1);
        agc2_init(&r->agc, 1e-1f, 1e-1f, 0.25f, 1.0f);
        int n3 = qo_firdes_band_pass_2(1, r->tsr, 200, filter_width, 200, 90, QO_WIN_BLACKMAN_HARRIS, T3, 4096);
        r->ntaps_store[3] = n3;
        resamp_init(&r->ssb_audio, 1, 1, 1, T3, n3);
        qv_init(&r->s_clip, 8);
        r->env_m2 = 0; r->env_m1 = 0; r->st_pos = 0;
        r->port[1].isz = 4;
    } else if (kind == QO_DEMOD_2FSK) {
        /* /root/reference/src/gr/gr_demod_2fsk.cpp:33-167 (fm variant; the band-filter variant is not restated yet) */
        int decim, interp = 1, nfilts;
        if (sps == 10) { r->tsr = 20000; r->sym_sps = sps; decim = 50; nfilts = 35 * r->sym_sps; }
        else if (sps >= 5) { r->tsr = 40000; r->sym_sps = sps * 2; decim = 25; nfilts = 35 * r->sym_sps; }
        else if (sps == 1) { r->tsr = 80000; r->sym_sps = 4; decim = 25; interp = 2; nfilts = 125 * r->sym_sps; }
        else { free(r); return NULL; }
        int spacing = flag ? 1 : 2;
        if ((nfilts % 2) == 0) nfilts += 1;
        int n0 = qo_firdes_low_pass(interp, (double)interp * samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, interp, decim, T0, n0);
        fll_init(&r->fll, (float)r->sym_sps, 0.1f, 16, (float)(24 * M_PI / 100));
        int n1 = qo_firdes_low_pass(1, r->tsr, filter_width, filter_width, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->filt, 2, 1, 1, T1, n1);
        qdemod_init(&r->qd, (float)(r->sym_sps / (spacing * M_PI / 2)));
        int n2 = gmsk ? qo_firdes_low_pass(1, r->tsr, r->tsr / r->sym_sps, r->tsr / r->sym_sps, QO_WIN_HAMMING, T2, 4096)   /* gr_demod_gmsk.cpp:88-90 */
                      : qo_firdes_rrc(1, r->tsr, r->tsr / r->sym_sps, 0.2, nfilts, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->shaping, 1, 1, 1, T2, n2);
        if (!flag) {
            /* band-filter variant (gr_demod_2fsk.cpp:88-100,137-149): upper = (-fw,0), lower = (0,fw), |upper|/|lower| */
            float tc[2 * 4096];
            int nb = qo_firdes_complex_band_pass(1, r->tsr, -filter_width, 0, filter_width, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
            fircc_init(&r->bp[0], tc, nb);                       /* _upper_filter */
            nb = qo_firdes_complex_band_pass(1, r->tsr, 0, filter_width, filter_width, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
            fircc_init(&r->bp[1], tc, nb);                       /* _lower_filter */
            int n3 = qo_firdes_low_pass(1.0, r->tsr, r->tsr / r->sym_sps, r->tsr / r->sym_sps, QO_WIN_HAMMING, T3, 4096);
            r->ntaps_store[3] = n3;
            resamp_init(&r->symfilt, 1, 1, 1, T3, n3);
        }
        float symbol_rate = (float)r->tsr / (float)r->sym_sps;
        float sps_dev = 200.0f / symbol_rate;
        if (gmsk) symsync_init(&r->ss, 1, (float)r->sym_sps, (float)(2 * M_PI / 200.0f), 1.0f, 0.2869f, 0.05f, SL_BPSK);   /* gr_demod_gmsk.cpp:80-84 */
        else symsync_init(&r->ss, 1, (float)r->sym_sps, (float)(2 * M_PI / (symbol_rate / 10)), 1.0f, 0.2869f, sps_dev, SL_BPSK);
        r->soft_scale = 128.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr); ccdec_init(&r->dec2); lfsr_init(&r->descr2);
        r->fm = flag;
    } else if (kind == QO_DEMOD_BPSK) {
        /* /root/reference/src/gr/gr_demod_bpsk.cpp:33-105 */
        r->tsr = 20000; r->sym_sps = sps;
        int n0 = qo_firdes_low_pass(1, samp_rate, r->tsr / 2, r->tsr / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, 50, T0, n0);
        fll_init(&r->fll, (float)sps, 0.35f, 32, (float)(8 * M_PI / 100));
        int n1 = qo_firdes_rrc(sps, sps, 1, 0.35, 15 * sps, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->shaping, 2, 1, 1, T1, n1);
        agc2_init(&r->agc, 1e-1f, 1e-1f, 1.0f, 1.0f);
        float gain_mu = 0.05f, gain_omega = 0.005f;
        crmm_init(&r->crmm, (float)sps, gain_omega * gain_omega, 0.5f, gain_mu, 0.001f);
        costas_init(&r->costas, (float)(2 * M_PI / 200), 2, 0);
        r->soft_scale = 64.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr); ccdec_init(&r->dec2); lfsr_init(&r->descr2);
    } else if (kind == QO_DEMOD_DSSS) {
        /* /root/reference/src/gr/gr_demod_dsss.cpp:32-124 (instance gr_demod_base.cpp:218: make_gr_demod_dsss(25, 1e6, 1700, 150)):
         * /50 -> rational_resampler_ccf(13, 50) low_pass(1, 20k, 2600, 2600, BH) -> costas_loop_cc(pi/200, 2, snr) -> low_pass(1, 5200,
         * fw, 1200, BH) [port 0] -> agc2_cc(.1, .1, 1, 10) -> dsss_decoder_cc(barker 13, sps) -> clock_recovery_mm_cc(1, 2.5e-5, .5,
         * .05, .005) -> costas_loop_cc(2 pi/100, 2) [port 1] -> real -> x64 + 128 -> uchar -> 2 x cc_decoder (second after delay(1))
         * -> 2 x descrambler [ports 2, 3] */
        static const int barker_13[13] = { 1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 0, 1 };
        r->tsr = 5200; r->sym_sps = sps;
        int n0 = qo_firdes_low_pass(1, samp_rate, 20000 / 2, 20000 / 2, QO_WIN_BLACKMAN_HARRIS, T0, 4096);
        r->ntaps_store[0] = n0;
        resamp_init(&r->resamp, 2, 1, 50, T0, n0);
        int n1 = qo_firdes_low_pass(1, 20000, 5200 / 2, 5200 / 2, QO_WIN_BLACKMAN_HARRIS, T1, 4096);
        r->ntaps_store[1] = n1;
        resamp_init(&r->resamp_if, 2, 13, 50, T1, n1);
        costas_init(&r->pll, (float)(M_PI / 200), 2, 1);
        int n2 = qo_firdes_low_pass(1, 5200, filter_width, 1200, QO_WIN_BLACKMAN_HARRIS, T2, 4096);
        r->ntaps_store[2] = n2;
        resamp_init(&r->shaping, 2, 1, 1, T2, n2);
        agc2_init(&r->agc, 1e-1f, 1e-1f, 1.0f, 10.0f);
        dsssdec_init(&r->dsss, barker_13, 13, sps);
        float gain_mu = 0.05f, gain_omega = 0.005f;
        crmm_init(&r->crmm, 1.0f, gain_omega * gain_omega, 0.5f, gain_mu, 0.005f);
        costas_init(&r->costas, (float)(2 * M_PI / 100), 2, 0);
        r->soft_scale = 64.0f;
        ccdec_init(&r->dec); lfsr_init(&r->descr); ccdec_init(&r->dec2); lfsr_init(&r->descr2);
        qv_init(&r->s_dsss, 8);
    } else { free(r); return NULL; }
    if (r->kind == QO_DEMOD_NBFM || r->kind == QO_DEMOD_WBFM) { r->filt.hkeep = 2048; r->audio_filt.hkeep = 512; }      /* filters a setter may lengthen */
    if (r->kind == QO_DEMOD_SSB || r->kind == QO_DEMOD_AM) { r->ssb_bpf.hkeep = 2048; }
    if (r->kind == QO_DEMOD_SSB) r->ssb_audio.hkeep = 512;
    return r;
}
void qo_rx_destroy(qo_rx* r)
{
    if (!r) return;
    /* leak-tolerant: test infrastructure; free the big ones */
    qv_free(&r->s_res); qv_free(&r->s_filt); qv_free(&r->s_dem); qv_free(&r->s_rrc); qv_free(&r->s_sym);
    qv_free(&r->s_soft); qv_free(&r->s_bits); qv_free(&r->s_tmp); qv_free(&r->s_tmp2);
    for (int i = 0; i < 4; i++) { qv_free(&r->port[i]); qv_free(&r->s_bp[i]); }
    free(r);
}

static void rx_fec_tail(qo_rx* r)
{
    /* soft -> decoder -> descrambler -> port2 */
    size_t b0 = r->s_bits.n;
    ccdec_work(&r->dec, r->s_soft.d, r->s_soft.n, &r->s_bits);
    r->s_soft.n = 0;
    for (size_t i = b0; i < r->s_bits.n; i++) qv_pushb(&r->port[2], lfsr_descramble(&r->descr, r->s_bits.d[i]));
    r->s_bits.n = 0;
}

/* two decoders on the same soft stream, the second one behind a delay(1) (gr_demod_bpsk.cpp:96-104) */
static void rx_fec_tail_dual(qo_rx* r)
{
    size_t b0 = r->s_bits.n;
    ccdec_work(&r->dec, r->s_soft.d, r->s_soft.n, &r->s_bits);
    for (size_t i = b0; i < r->s_bits.n; i++) qv_pushb(&r->port[2], lfsr_descramble(&r->descr, r->s_bits.d[i]));
    r->s_bits.n = 0;
    if (!r->dec2_started) { unsigned char z = 0; ccdec_work(&r->dec2, &z, 1, &r->s_bits); r->dec2_started = 1; }
    ccdec_work(&r->dec2, r->s_soft.d, r->s_soft.n, &r->s_bits);
    for (size_t i = 0; i < r->s_bits.n; i++) qv_pushb(&r->port[3], lfsr_descramble(&r->descr2, r->s_bits.d[i]));
    r->s_bits.n = 0;
    r->s_soft.n = 0;
}

/* blocks::rotator_cc restated with an exact Q32 phase (inc = rint(-offset/fs * 2^32)): y[n] = x[n] * exp(j*theta_n),
 * theta_n = phase_n * pi / 2^31 (phase as signed 32-bit), sin/cos from qo_sincosf.  GNU Radio's rotator accumulates a
 * float complex phasor (and renormalises every 512 samples); the two agree to ~1e-6. */
/* Run-time setters of the analog blocks (gr_demod_nbfm.cpp:82-121, gr_demod_ssb.cpp:89-121, gr_demod_am.cpp:84-107,
 * gr_demod_wbfm.cpp:77-91).  New taps meet the true sample history from the next output on.  Returns 0, or -1 when the block has
 * no such setter. */
int qo_rx_set_param(qo_rx* r, int key, double value)
{
    static float T[4096]; static float TC[2 * 4096];
    const int analog = r->kind == QO_DEMOD_NBFM || r->kind == QO_DEMOD_SSB || r->kind == QO_DEMOD_AM || r->kind == QO_DEMOD_WBFM;
    if (!analog) return -1;
    if (key == QO_PARAM_SQUELCH_DB) { r->sq.threshold = pow(10.0, value / 10.0); return 0; }        /* squelch_base::set_threshold */
    if (key == QO_PARAM_AGC_ATTACK && (r->kind == QO_DEMOD_SSB || r->kind == QO_DEMOD_AM)) { r->agc.attack = (float)value; return 0; }
    if (key == QO_PARAM_AGC_DECAY && (r->kind == QO_DEMOD_SSB || r->kind == QO_DEMOD_AM)) { r->agc.decay = (float)value; return 0; }
    if (key == QO_PARAM_GAIN && r->kind == QO_DEMOD_SSB) { r->if_gain = (float)value; return 0; }       /* _if_gain->set_k */
    if (key == QO_PARAM_CTCSS && r->kind == QO_DEMOD_NBFM && value == 0.0) {
        /* gr_demod_nbfm.cpp:99-111: the first disconnect throws when the tone squelch is not in the graph and the rest is skipped */
        if (r->ctcss_on) {
            int n3 = qo_firdes_low_pass_2(1, 8000, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, T, 4096);
            resamp_retap(&r->audio_filt, T, n3);
            r->ctcss_on = 0;
        }
        return 0;
    }
    if (key == QO_PARAM_CTCSS && r->kind == QO_DEMOD_NBFM) {
        /* :112-125: set_frequency always (new Goertzel filters, block state kept); graph and audio filter only on the first switch */
        ctcss_set_frequency(&r->ctcss, (float)value);
        if (!r->ctcss_on) {
            int n3 = qo_firdes_band_pass_2(1, 8000, 300, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, T, 4096);
            resamp_retap(&r->audio_filt, T, n3);
            r->ctcss_on = 1;
        }
        return 0;
    }
    if (key == QO_PARAM_FILTER_WIDTH) {
        const int fw = (int)value;
        if (fw <= 0) return -1;
        r->filter_width = fw;
        if (r->kind == QO_DEMOD_NBFM) {
            int n = qo_firdes_low_pass(1, r->tsr, fw, 1200, QO_WIN_BLACKMAN_HARRIS, T, 4096);
            resamp_retap(&r->filt, T, n);
            r->qd.gain = (float)(r->tsr / (4 * M_PI * fw));
        } else if (r->kind == QO_DEMOD_WBFM) {
            int n = qo_firdes_low_pass(1, r->tsr, fw, 1200, QO_WIN_BLACKMAN_HARRIS, T, 4096);
            resamp_retap(&r->filt, T, n);
            r->qd.gain = (float)(r->tsr / (2 * M_PI * fw));
        } else if (r->kind == QO_DEMOD_AM) {
            int n = qo_firdes_complex_band_pass(1, r->tsr, -fw, fw, 1200, QO_WIN_BLACKMAN_HARRIS, TC, 4096);
            fircc_retap(&r->ssb_bpf, TC, n);
        } else {
            int n = r->flag ? qo_firdes_complex_band_pass_2(1, r->tsr, -fw, -200, 200, 90, QO_WIN_BLACKMAN_HARRIS, TC, 4096)
                            : qo_firdes_complex_band_pass_2(1, r->tsr, 200, fw, 200, 90, QO_WIN_BLACKMAN_HARRIS, TC, 4096);
            fircc_retap(&r->ssb_bpf, TC, n);
            int n3 = qo_firdes_band_pass_2(2, r->tsr, 200, fw, 200, 90, QO_WIN_BLACKMAN_HARRIS, T, 4096);      /* gain 2 here, 1 in the constructor: the reference's */
            resamp_retap(&r->ssb_audio, T, n3);
        }
        return 0;
    }
    return -1;
}

void qo_rx_set_carrier_offset(qo_rx* r, double offset_hz, double samp_rate)
{
    r->rot_base = r->rot_base + r->rot_inc * (uint32_t)(r->rot_n - r->rot_nbase);
    r->rot_nbase = r->rot_n;
    r->rot_inc = (uint32_t)(int32_t)(long long)rint(-offset_hz / samp_rate * 4294967296.0);
}
static const float* rx_rotate(qo_rx* r, const float* iq, long T)
{
    if (r->rot_inc == 0 && r->rot_base == 0) { r->rot_n += T; return iq; }
    if (!r->s_rot.isz) qv_init(&r->s_rot, 8);
    r->s_rot.n = 0;
    for (long i = 0; i < T; i++) {
        uint32_t ph = r->rot_base + r->rot_inc * (uint32_t)(r->rot_n + i - r->rot_nbase);
        float ang = (float)((double)(int32_t)ph * (M_PI / 2147483648.0));
        float sn, cs; qo_sincosf(ang, &sn, &cs);
        float xr = iq[2 * i], xi = iq[2 * i + 1];
        qv_pushc(&r->s_rot, xr * cs - xi * sn, xr * sn + xi * cs);
    }
    r->rot_n += T;
    return (const float*)r->s_rot.d;
}


/* ------------------------------------------------------------------ gr_demod_base front end at device rates >= 2 Msps
 * /root/reference/src/gr/gr_demod_base.cpp:57-63,180,1220-1225,1303-1362: source -> rotator_cc(2 pi (-offset) / samp_rate) ->
 * rational_resampler_ccf(1, samp_rate / 1e6, low_pass(1, samp_rate, 480000, 100000, BLACKMAN_HARRIS)) -> demodulators at 1 Msps.
 * (Below 2 Msps the resampler is not in the graph and the rotator feeds the demodulators directly: that case is qo_rx's own rotator.)
 * Same Q32 NCO as rx_rotate, same FIR order as every decimator here. */
struct qo_frontend { int samp_rate, D; resamp_t rs; uint32_t inc, base; long long nbase, n; qvec s_rot; };
typedef struct qo_frontend qo_frontend;
qo_frontend* qo_frontend_create(int samp_rate)
{
    if (samp_rate < 2000000 || samp_rate % 1000000) return NULL;
    qo_frontend* f = (qo_frontend*)calloc(1, sizeof *f);
    f->samp_rate = samp_rate; f->D = samp_rate / 1000000;
    static float T[8192];
    int n = qo_firdes_low_pass(1, samp_rate, 480000, 100000, QO_WIN_BLACKMAN_HARRIS, T, 8192);
    resamp_init(&f->rs, 2, 1, f->D, T, n);
    qv_init(&f->s_rot, 8);
    return f;
}
void qo_frontend_destroy(qo_frontend* f) { if (f) { resamp_free(&f->rs); qv_free(&f->s_rot); free(f); } }
int qo_frontend_ntaps(const qo_frontend* f) { return f->rs.nt; }
void qo_frontend_set_carrier_offset(qo_frontend* f, double offset_hz)
{
    f->base = f->base + f->inc * (uint32_t)(f->n - f->nbase);
    f->nbase = f->n;
    f->inc = (uint32_t)(int32_t)(long long)rint(-offset_hz / f->samp_rate * 4294967296.0);
}
/* n input samples at the device rate -> returns the 1 Msps samples written to out (complex interleaved, cap items) */
long qo_frontend_work(qo_frontend* f, const float* iq, long n, float* out, long cap)
{
    const float* x = iq;
    if (f->inc != 0 || f->base != 0) {
        f->s_rot.n = 0;
        for (long i = 0; i < n; i++) {
            uint32_t ph = f->base + f->inc * (uint32_t)(f->n + i - f->nbase);
            float ang = (float)((double)(int32_t)ph * (M_PI / 2147483648.0));
            float sn, cs; qo_sincosf(ang, &sn, &cs);
            float xr = iq[2 * i], xi = iq[2 * i + 1];
            qv_pushc(&f->s_rot, xr * cs - xi * sn, xr * sn + xi * cs);
        }
        x = (const float*)f->s_rot.d;
    }
    f->n += n;
    qvec o; qv_init(&o, 8);
    resamp_work(&f->rs, x, (size_t)n, &o);
    long m = (long)o.n < cap ? (long)o.n : cap;
    memcpy(out, o.d, (size_t)m * 8);
    qv_free(&o);
    return m;
}

/* ---- in-tree reference blocks restated as single-item functions (pinned bit-for-bit against the reference sources compiled
 *      in oracle/_ref: tests/test_oracle_ref.py) ---- */
/* cessb::clipper_cc (cessb/clipper_cc_impl.cc:65-95): magnitude limited to `clip`, phase kept (fast_atan2f, then cos / sin) */
static void cessb_clip_one(float re, float im, float clip, float* orr, float* oi)
{
    float mag = sqrtf(re * re + im * im);
    float ph = qo_fast_atan2f(im, re);
    float cl = mag < clip ? mag : clip;
    float sn, cs; qo_sincosf(ph, &sn, &cs);
    *orr = cs * cl; *oi = sn * cl;
}
/* cessb::stretcher_cc (cessb/stretcher_cc_impl.cc:70-110): divisor from the 5-point envelope maximum around the item */
static float cessb_stretch_div(float e_m2, float e_m1, float e0, float e1, float e2)
{
    const float emax = (float)(1 / (sqrt(0.5) / 2));
    float h = e0;
    h = fmaxf(h, e_m2); h = fmaxf(h, e_m1); h = fmaxf(h, e1); h = fmaxf(h, e2);
    h = h * emax; h = fmaxf(h, 1.0f); h = h - 1.0f; h = h * 2.0f; h = h + 1.0f;
    return h;
}
/* gr_4fsk_discriminator::work (gr_4fsk_discriminator.cpp:17-44): strict-greater argmax of four magnitudes */
static void disc4_one(const float* m, float* orr, float* oi)
{
    *orr = 0; *oi = 0;
    if ((m[0] > m[1]) && (m[0] > m[2]) && (m[0] > m[3])) { *orr = (float)-0.707107; *oi = (float)-0.707107; }
    else if ((m[1] > m[0]) && (m[1] > m[2]) && (m[1] > m[3])) { *orr = (float)-0.707107; *oi = (float)0.707107; }
    else if ((m[2] > m[1]) && (m[2] > m[0]) && (m[2] > m[3])) { *orr = (float)0.707107; *oi = (float)0.707107; }
    else if ((m[3] > m[1]) && (m[3] > m[0]) && (m[3] > m[2])) { *orr = (float)0.707107; *oi = (float)-0.707107; }
}
void qo_cessb_clipper(const float* in_c, long n, float clip, float* out_c)
{
    for (long i = 0; i < n; i++) cessb_clip_one(in_c[2 * i], in_c[2 * i + 1], clip, &out_c[2 * i], &out_c[2 * i + 1]);
}
/* one-shot stretcher over a stream that starts at item 0 (zero envelope history): writes n - 2 items, returns that count */
long qo_cessb_stretcher(const float* in_c, long n, float* out_c)
{
    float em2 = 0.0f, em1 = 0.0f;
    long k = 0;
    for (; k + 2 < n; k++) {
        const float* c = in_c + 2 * k;
        float e0 = sqrtf(c[0] * c[0] + c[1] * c[1]), e1 = sqrtf(c[2] * c[2] + c[3] * c[3]), e2 = sqrtf(c[4] * c[4] + c[5] * c[5]);
        float h = cessb_stretch_div(em2, em1, e0, e1, e2);
        out_c[2 * k] = c[0] / h; out_c[2 * k + 1] = c[1] / h;
        em2 = em1; em1 = e0;
    }
    return k;
}
void qo_disc4(const float* m0, const float* m1, const float* m2, const float* m3, long n, float* out_c)
{
    for (long i = 0; i < n; i++) { float m[4] = { m0[i], m1[i], m2[i], m3[i] }; disc4_one(m, &out_c[2 * i], &out_c[2 * i + 1]); }
}

int qo_rx_work(qo_rx* r, const float* iq, long T)
{
    iq = rx_rotate(r, iq, T);
    if (r->kind == QO_DEMOD_SSB) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        float* v = (float*)r->s_res.d;
        for (size_t i = 0; i < 2 * r->s_res.n; i++) v[i] = v[i] * r->if_gain;           /* multiply_const_cc(0.9) (set_gain) */
        r->s_filt.n = 0; fircc_work(&r->ssb_bpf, v, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_tmp.n = 0; squelch_work(&r->sq, (const float*)r->s_filt.d, r->s_filt.n, &r->s_tmp);
        /* agc2_cc -> cessb::clipper_cc(0.95) (cessb/clipper_cc_impl.cc:65-95): magnitude clipped, phase kept */
        const float* g = (const float*)r->s_tmp.d;
        for (size_t i = 0; i < r->s_tmp.n; i++) {
            float ar, ai;
            agc2_step(&r->agc, g[2 * i], g[2 * i + 1], &ar, &ai);
            float cr, ci;
            cessb_clip_one(ar, ai, 0.95f, &cr, &ci);
            qv_pushc(&r->s_clip, cr, ci);
        }
        /* cessb::stretcher_cc (stretcher_cc_impl.cc:70-110): 5-point envelope hold with a 2-sample look-ahead */
        const float* c = (const float*)r->s_clip.d;
        r->s_rrc.n = 0;
        while (r->st_pos + 2 < r->s_clip.n) {
            size_t n = r->st_pos;
            float e0 = sqrtf(c[2 * n] * c[2 * n] + c[2 * n + 1] * c[2 * n + 1]);
            float e1 = sqrtf(c[2 * (n + 1)] * c[2 * (n + 1)] + c[2 * (n + 1) + 1] * c[2 * (n + 1) + 1]);
            float e2 = sqrtf(c[2 * (n + 2)] * c[2 * (n + 2)] + c[2 * (n + 2) + 1] * c[2 * (n + 2) + 1]);
            float h = cessb_stretch_div(r->env_m2, r->env_m1, e0, e1, e2);
            float re = c[2 * n] / h;
            qv_pushf(&r->s_rrc, re * 1.333f);                                            /* complex_to_real, x1.333 */
            r->env_m2 = r->env_m1; r->env_m1 = e0;
            r->st_pos++;
        }
        if (r->st_pos > 0) { qv_drop(&r->s_clip, r->st_pos); r->st_pos = 0; }
        resamp_work(&r->ssb_audio, (const float*)r->s_rrc.d, r->s_rrc.n, &r->port[1]);
        return 0;
    }
    if (r->kind == QO_DEMOD_2FSK) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        float* v = (float*)r->s_res.d;
        if (!r->gmsk) for (size_t i = 0; i < r->s_res.n; i++) fll_step(&r->fll, v[2 * i], v[2 * i + 1], &v[2 * i], &v[2 * i + 1]);
        r->s_filt.n = 0; resamp_work(&r->filt, v, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_rrc.n = 0;
        if (r->fm) {
            r->s_dem.n = 0; qdemod_work(&r->qd, (const float*)r->s_filt.d, r->s_filt.n, &r->s_dem);
            resamp_work(&r->shaping, (const float*)r->s_dem.d, r->s_dem.n, &r->s_rrc);
        } else {
            const float* f = (const float*)r->s_filt.d; size_t n = r->s_filt.n;
            for (int k = 0; k < 2; k++) { r->s_bp[k].n = 0; fircc_work(&r->bp[k], f, n, &r->s_bp[k]); }
            r->s_dem.n = 0;
            for (size_t i = 0; i < n; i++) {
                const float* u = (const float*)r->s_bp[0].d + 2 * i; const float* l = (const float*)r->s_bp[1].d + 2 * i;
                float mu_ = sqrtf(u[0] * u[0] + u[1] * u[1]), ml = sqrtf(l[0] * l[0] + l[1] * l[1]);
                float q = mu_ / ml;                                   /* blocks::divide_ff: upper / lower */
                /* analog::rail_ff(0, 2).  0/0 at start-up gives NaN; the clamp is restated with IEEE fminf/fmaxf
                 * (NaN -> the rail), which is what a SIMD min/max clamp does; std::min/max would let the NaN through
                 * and poison the clock loop for ever. */
                float t = fmaxf(fminf(q, 2.0f), 0.0f);
                qv_pushf(&r->s_dem, t + -1.0f);                       /* add_const_ff(-1) */
            }
            resamp_work(&r->symfilt, (const float*)r->s_dem.d, r->s_dem.n, &r->s_rrc);
        }
        r->s_sym.isz = 4; r->s_sym.n = 0;
        symsync_work(&r->ss, (const float*)r->s_rrc.d, r->s_rrc.n, &r->s_sym);
        const float* sy = (const float*)r->s_sym.d;
        for (size_t i = 0; i < r->s_sym.n; i++) {
            qv_pushc(&r->port[1], sy[i], 0.0f);                      /* float_to_complex, imag 0 */
            qv_pushb(&r->s_soft, soft_u8(sy[i], r->soft_scale));
        }
        rx_fec_tail_dual(r);
        return 0;
    }
    if (r->kind == QO_DEMOD_DSSS) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        r->s_tmp.n = 0; resamp_work(&r->resamp_if, (const float*)r->s_res.d, r->s_res.n, &r->s_tmp);
        float* v = (float*)r->s_tmp.d;
        for (size_t i = 0; i < r->s_tmp.n; i++) costas_step(&r->pll, v[2 * i], v[2 * i + 1], &v[2 * i], &v[2 * i + 1]);
        r->s_filt.n = 0; resamp_work(&r->shaping, v, r->s_tmp.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        float* f = (float*)r->s_filt.d;
        for (size_t i = 0; i < r->s_filt.n; i++) agc2_step(&r->agc, f[2 * i], f[2 * i + 1], &f[2 * i], &f[2 * i + 1]);
        r->s_dsss.n = 0; dsssdec_work(&r->dsss, f, r->s_filt.n, &r->s_dsss);
        r->s_sym.isz = 8; r->s_sym.n = 0;
        crmm_work(&r->crmm, (const float*)r->s_dsss.d, r->s_dsss.n, &r->s_sym);
        const float* sy = (const float*)r->s_sym.d;
        for (size_t i = 0; i < r->s_sym.n; i++) {
            float cr, ci;
            costas_step(&r->costas, sy[2 * i], sy[2 * i + 1], &cr, &ci);
            qv_pushc(&r->port[1], cr, ci);
            qv_pushb(&r->s_soft, soft_u8(cr, r->soft_scale));
        }
        rx_fec_tail_dual(r);
        return 0;
    }
    if (r->kind == QO_DEMOD_BPSK) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        float* v = (float*)r->s_res.d;
        for (size_t i = 0; i < r->s_res.n; i++) fll_step(&r->fll, v[2 * i], v[2 * i + 1], &v[2 * i], &v[2 * i + 1]);
        r->s_filt.n = 0; resamp_work(&r->shaping, v, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        float* f = (float*)r->s_filt.d;
        for (size_t i = 0; i < r->s_filt.n; i++) agc2_step(&r->agc, f[2 * i], f[2 * i + 1], &f[2 * i], &f[2 * i + 1]);
        r->s_sym.isz = 8; r->s_sym.n = 0;
        crmm_work(&r->crmm, f, r->s_filt.n, &r->s_sym);
        const float* sy = (const float*)r->s_sym.d;
        for (size_t i = 0; i < r->s_sym.n; i++) {
            float cr, ci;
            costas_step(&r->costas, sy[2 * i], sy[2 * i + 1], &cr, &ci);
            qv_pushc(&r->port[1], cr, ci);
            qv_pushb(&r->s_soft, soft_u8(cr, r->soft_scale));      /* complex_to_real */
        }
        rx_fec_tail_dual(r);
        return 0;
    }
    if (r->kind == QO_DEMOD_4FSK) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        r->s_filt.n = 0; resamp_work(&r->filt, (const float*)r->s_res.d, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_sym.n = 0;
        if (r->fm) {
            r->s_dem.n = 0; qdemod_work(&r->qd, (const float*)r->s_filt.d, r->s_filt.n, &r->s_dem);
            r->s_rrc.n = 0; resamp_work(&r->shaping, (const float*)r->s_dem.d, r->s_dem.n, &r->s_rrc);
            r->s_sym.isz = 4;
            symsync_work(&r->ss, (const float*)r->s_rrc.d, r->s_rrc.n, &r->s_sym);
            const float* sy = (const float*)r->s_sym.d;
            for (size_t i = 0; i < r->s_sym.n; i++) {
                /* analog::phase_modulator_fc(pi/2) */
                float ph = r->pm_sens * sy[i];
                float sn, cs; qo_sincosf(ph, &sn, &cs);
                qv_pushc(&r->port[1], cs, sn);
                if (r->m17) {
                    /* gr_demod_m17.cpp:98-107: interleave (real, imag) -> binary_slicer_fb (x >= 0) -> pack_k_bits(2) ->
                     * map {3,1,2,0} -> unpack_k_bits(2) */
                    static const int map[4] = { 3, 1, 2, 0 };
                    const int v = ((cs >= 0.0f) ? 2 : 0) | ((sn >= 0.0f) ? 1 : 0);
                    qv_pushb(&r->port[2], (unsigned char)((map[v] >> 1) & 1));
                    qv_pushb(&r->port[2], (unsigned char)(map[v] & 1));
                    continue;
                }
                /* interleave: imag first, then real (gr_demod_4fsk.cpp:186-189) */
                qv_pushb(&r->s_soft, soft_u8(sn, r->soft_scale));
                qv_pushb(&r->s_soft, soft_u8(cs, r->soft_scale));
            }
        } else {
            const float* f = (const float*)r->s_filt.d; size_t n = r->s_filt.n;
            for (int i = 0; i < 4; i++) { r->s_bp[i].n = 0; fircc_work(&r->bp[i], f, n, &r->s_bp[i]); }
            r->s_tmp.n = 0;
            for (size_t i = 0; i < n; i++) {
                float m[4];
                for (int k = 0; k < 4; k++) { const float* c = (const float*)r->s_bp[k].d + 2 * i; m[k] = sqrtf(c[0] * c[0] + c[1] * c[1]); }
                /* /root/reference/src/gr/gr_4fsk_discriminator.cpp:17-44 */
                float orr, oi;
                disc4_one(m, &orr, &oi);
                qv_pushc(&r->s_tmp, orr, oi);
            }
            r->s_tmp2.n = 0; resamp_work(&r->symfilt, (const float*)r->s_tmp.d, r->s_tmp.n, &r->s_tmp2);
            r->s_sym.isz = 8;
            symsync_work(&r->ss, (const float*)r->s_tmp2.d, r->s_tmp2.n, &r->s_sym);
            const float* sy = (const float*)r->s_sym.d;
            for (size_t i = 0; i < r->s_sym.n; i++) {
                qv_pushc(&r->port[1], sy[2 * i], sy[2 * i + 1]);
                qv_pushb(&r->s_soft, soft_u8(sy[2 * i], r->soft_scale));
                qv_pushb(&r->s_soft, soft_u8(sy[2 * i + 1], r->soft_scale));
            }
        }
        if (!r->m17) rx_fec_tail(r);
        return 0;
    }
    if (r->kind == QO_DEMOD_QPSK) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        if (r->fm) {
            float* v = (float*)r->s_res.d;
            for (size_t i = 0; i < r->s_res.n; i++) fll_step(&r->fll, v[2 * i], v[2 * i + 1], &v[2 * i], &v[2 * i + 1]);
        }
        r->s_filt.n = 0; resamp_work(&r->shaping, (const float*)r->s_res.d, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        const float* f = (const float*)r->s_filt.d; size_t n = r->s_filt.n;
        r->s_tmp.n = 0;
        for (size_t i = 0; i < n; i++) {
            float ar, ai, pr, pi;
            agc2_step(&r->agc, f[2 * i], f[2 * i + 1], &ar, &ai);
            costas_step(&r->pll, ar, ai, &pr, &pi);
            qv_pushc(&r->s_tmp, pr, pi);
        }
        r->s_sym.isz = 8; r->s_sym.n = 0;
        symsync_work(&r->ss, (const float*)r->s_tmp.d, r->s_tmp.n, &r->s_sym);
        const float* sy = (const float*)r->s_sym.d;
        for (size_t i = 0; i < r->s_sym.n; i++) {
            float cr, ci;
            costas_step(&r->costas, sy[2 * i], sy[2 * i + 1], &cr, &ci);
            /* digital::diff_phasor_cc */
            float dr = cr * r->dp_r + ci * r->dp_i;
            float di = ci * r->dp_r - cr * r->dp_i;
            r->dp_r = cr; r->dp_i = ci;
            /* blocks::multiply_const_cc(exp(-j 3pi/4)) */
            float orr = dr * r->rot_r - di * r->rot_i;
            float oi = dr * r->rot_i + di * r->rot_r;
            qv_pushc(&r->port[1], orr, oi);
            qv_pushb(&r->s_soft, soft_u8(orr, r->soft_scale));
            qv_pushb(&r->s_soft, soft_u8(oi, r->soft_scale));
        }
        rx_fec_tail(r);
        return 0;
    }
    if (r->kind == QO_DEMOD_DMR) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        qv_push(&r->port[0], r->s_res.d, r->s_res.n);
        r->s_dem.n = 0; qdemod_work(&r->qd, (const float*)r->s_res.d, r->s_res.n, &r->s_dem);
        r->s_rrc.n = 0; resamp_work(&r->shaping, (const float*)r->s_dem.d, r->s_dem.n, &r->s_rrc);
        qv_push(&r->port[3], r->s_rrc.d, r->s_rrc.n);
        r->s_sym.isz = 4; r->s_sym.n = 0;
        symsync_work(&r->ss, (const float*)r->s_rrc.d, r->s_rrc.n, &r->s_sym);
        const float* sy = (const float*)r->s_sym.d;
        static const int map[4] = { 3, 1, 2, 0 };
        for (size_t i = 0; i < r->s_sym.n; i++) {
            const float ph = r->pm_sens * (sy[i] * 0.9f);                 /* multiply_const_ff(0.9) -> phase_modulator_fc(pi/2) */
            float sn, cs; qo_sincosf(ph, &sn, &cs);
            qv_pushc(&r->port[1], cs, sn);
            const int v = ((cs >= 0.0f) ? 2 : 0) | ((sn >= 0.0f) ? 1 : 0);
            qv_pushb(&r->port[2], (unsigned char)((map[v] >> 1) & 1));
            qv_pushb(&r->port[2], (unsigned char)(map[v] & 1));
        }
        return 0;
    }
    if (r->kind == QO_DEMOD_WBFM) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        r->s_filt.n = 0; resamp_work(&r->filt, (const float*)r->s_res.d, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_tmp.n = 0; squelch_work(&r->sq, (const float*)r->s_filt.d, r->s_filt.n, &r->s_tmp);
        r->s_dem.n = 0; qdemod_work(&r->qd, (const float*)r->s_tmp.d, r->s_tmp.n, &r->s_dem);
        float* d = (float*)r->s_dem.d;
        for (size_t i = 0; i < r->s_dem.n; i++) d[i] = d[i] * 0.9f;                        /* multiply_const_ff(0.9) */
        r->s_sym.isz = 4; r->s_sym.n = 0; iir1_work(&r->deemph, d, r->s_dem.n, &r->s_sym, 1.0f);
        resamp_work(&r->audio_rs, (const float*)r->s_sym.d, r->s_sym.n, &r->port[1]);
        return 0;
    }
    if (r->kind == QO_DEMOD_AM) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        r->s_filt.n = 0; fircc_work(&r->ssb_bpf, (const float*)r->s_res.d, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_tmp.n = 0; squelch_work(&r->sq, (const float*)r->s_filt.d, r->s_filt.n, &r->s_tmp);
        const float* g = (const float*)r->s_tmp.d;
        r->s_dem.n = 0;
        for (size_t i = 0; i < r->s_tmp.n; i++) {
            /* complex_to_mag, then analog::kernel::agc2_ff::scale */
            const float mag = sqrtf(g[2 * i] * g[2 * i] + g[2 * i + 1] * g[2 * i + 1]);
            const float out = mag * r->agc.gain;
            const float tmp = fabsf(out) - r->agc.ref;
            float rate = r->agc.decay;
            if (fabsf(tmp) > r->agc.gain) rate = r->agc.attack;
            r->agc.gain -= tmp * rate;
            if (r->agc.gain < 0.0f) r->agc.gain = 10e-5f;
            if (r->agc.max_gain > 0.0f && r->agc.gain > r->agc.max_gain) r->agc.gain = r->agc.max_gain;
            qv_pushf(&r->s_dem, out);
        }
        r->s_sym.isz = 4; r->s_sym.n = 0; iir1_work(&r->deemph, (const float*)r->s_dem.d, r->s_dem.n, &r->s_sym, 0.99f);
        r->s_rrc.n = 0; resamp_work(&r->audio_rs, (const float*)r->s_sym.d, r->s_sym.n, &r->s_rrc);
        resamp_work(&r->audio_filt, (const float*)r->s_rrc.d, r->s_rrc.n, &r->port[1]);
        return 0;
    }
    if (r->kind == QO_DEMOD_NBFM) {
        r->s_res.n = 0; resamp_work(&r->resamp, iq, (size_t)T, &r->s_res);
        r->s_filt.n = 0; resamp_work(&r->filt, (const float*)r->s_res.d, r->s_res.n, &r->s_filt);
        qv_push(&r->port[0], r->s_filt.d, r->s_filt.n);
        r->s_tmp.n = 0; squelch_work(&r->sq, (const float*)r->s_filt.d, r->s_filt.n, &r->s_tmp);
        r->s_dem.n = 0; qdemod_work(&r->qd, (const float*)r->s_tmp.d, r->s_tmp.n, &r->s_dem);
        r->s_rrc.n = 0; resamp_work(&r->audio_rs, (const float*)r->s_dem.d, r->s_dem.n, &r->s_rrc);
        if (r->ctcss_on && r->kind == QO_DEMOD_NBFM) {
            r->s_tmp2.isz = 4; r->s_tmp2.n = 0; ctcss_work(&r->ctcss, (const float*)r->s_rrc.d, r->s_rrc.n, &r->s_tmp2);
            r->s_sym.isz = 4; r->s_sym.n = 0; resamp_work(&r->audio_filt, (const float*)r->s_tmp2.d, r->s_tmp2.n, &r->s_sym);
        } else {
            r->s_sym.isz = 4; r->s_sym.n = 0; resamp_work(&r->audio_filt, (const float*)r->s_rrc.d, r->s_rrc.n, &r->s_sym);
        }
        iir1_work(&r->deemph, (const float*)r->s_sym.d, r->s_sym.n, &r->port[1], 2.0f);
        return 0;
    }
    return -1;
}
long qo_rx_port_items(qo_rx* r, int port) { return (long)r->port[port].n; }
const void* qo_rx_port_data(qo_rx* r, int port) { return r->port[port].d; }
void qo_rx_port_clear(qo_rx* r, int port) { r->port[port].n = 0; }
static qvec* rx_dbg(qo_rx* r, const char* name)
{
    if (!strcmp(name, "resamp")) return &r->s_res;
    if (!strcmp(name, "filt")) return &r->s_filt;
    if (!strcmp(name, "demod")) return &r->s_dem;
    if (!strcmp(name, "rrc")) return &r->s_rrc;
    if (!strcmp(name, "sym")) return &r->s_sym;
    if (!strcmp(name, "tmp")) return &r->s_tmp;
    return NULL;
}
long qo_rx_dbg_items(qo_rx* r, const char* name) { qvec* v = rx_dbg(r, name); return v ? (long)v->n : -1; }
const void* qo_rx_dbg_data(qo_rx* r, const char* name) { qvec* v = rx_dbg(r, name); return v ? v->d : NULL; }
int qo_rx_ntaps(qo_rx* r, int which, float* out, int cap)
{
    int n = r->ntaps_store[which];
    if (out && n <= cap) memcpy(out, r->taps_store[which], sizeof(float) * n);
    return n;
}

/* ------------------------------------------------------------------ TX chains */
struct qo_tx {
    int kind, fm, sps;
    lfsr_t scr; ccenc_t enc;
    resamp_t rrc; resamp_t interp;
    float fm_sens, phase_f; uint32_t phase_q;
    float amplif, bb_gain;
    int pack_have; unsigned pack_acc;
    unsigned diff_prev;
    /* analog modulators */
    resamp_t a_filt, a_rs, a_if; iir1_t preemph; fircc_t a_sb; float env_m2, env_m1; size_t st_pos; qvec s_aud, s_clip, s_c2;
    qvec s_bits, s_coded, s_sym, s_shaped, s_mod, out;
    /* gr_mod_dmr: gr_zero_idle_bursts (a delay line of history-1 items + the "zero_samples" tags) */
    int dsss;
    int samp_rate, flag;
    float audio_gain; int tone_on; uint32_t tone_inc, tone_phase;      /* gr_mod_nbfm: multiply_const_ff and the CTCSS tone source */
    agc2_t am_agc; float am_dc;      /* gr_mod_am */
    int dmr; float* zi_line; long zi_len; unsigned zi_delay; uint64_t zi_n, zi_counter;
    long long* zi_tag_off; uint64_t* zi_tag_val; long zi_ntags, zi_cap;
};
qo_tx* qo_tx_create(int kind, int sps, int samp_rate, int carrier_freq, int filter_width, int flag)
{
    (void)carrier_freq;
    tabs_init();
    qo_tx* t = (qo_tx*)calloc(1, sizeof *t);
    t->kind = kind; t->bb_gain = 1.0f;
    lfsr_init(&t->scr);
    qv_init(&t->s_bits, 1); qv_init(&t->s_coded, 1); qv_init(&t->s_sym, 4); qv_init(&t->s_shaped, 4);
    qv_init(&t->s_mod, 8); qv_init(&t->out, 8);
    static float taps[16384];
    if (kind == QO_MOD_4FSK) {
        /* /root/reference/src/gr/gr_mod_4fsk.cpp:27-117 */
        int fm = flag; t->fm = fm;
        int sym_sps = sps, nfilts = sym_sps * 10, second_interp = 20;
        if (sps == 2) { sym_sps = 5; second_interp = 2; nfilts = 256; }
        int spacing = 2; t->amplif = 0.8f;
        if (fm) { t->amplif = 0.9f; spacing = 1; }
        t->sps = sym_sps;
        int n = qo_firdes_rrc(sym_sps, sym_sps, 1, 0.2, nfilts, taps, 16384);
        resamp_init(&t->rrc, 1, sym_sps, 1, taps, n);
        t->fm_sens = (float)((spacing * M_PI) / sym_sps);
        n = qo_firdes_low_pass(second_interp, samp_rate, filter_width, filter_width, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->interp, 2, second_interp, 1, taps, n);
        t->s_sym.isz = 4;
    } else if (kind == QO_MOD_M17) {
        /* /root/reference/src/gr/gr_mod_m17.cpp:30-95: bits (no scrambler / FEC in this block) -> pack 2 -> map {2,3,1,0} -> levels
         * -> rational_resampler_fff(5, 1, RRC(5, 5, 1, 0.5, 250)) -> x0.66666666 -> frequency_modulator_fc(pi / 5) -> low_pass(1, 24k,
         * fw, fw, BH) -> x0.9 -> x bb_gain -> rational_resampler_ccf(sps = 125, 3, low_pass(125, 3 fs, 12k, 12k, BH)) */
        t->fm = 1; t->sps = 5; t->amplif = 0.9f;
        int n = qo_firdes_rrc(5, 5, 1, 0.5, 250, taps, 16384);
        resamp_init(&t->rrc, 1, 5, 1, taps, n);
        t->fm_sens = (float)(M_PI / 5);
        n = qo_firdes_low_pass(1, 24000, filter_width, filter_width, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->a_if, 2, 1, 1, taps, n);
        n = qo_firdes_low_pass(sps, 3.0 * samp_rate, 12000, 12000, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->interp, 2, sps, 3, taps, n);
        t->s_sym.isz = 4; qv_init(&t->s_c2, 8);
    } else if (kind == QO_MOD_DMR) {
        /* /root/reference/src/gr/gr_mod_dmr.cpp:27-93: the gr_mod_m17 bit chain (pack 2 -> map {2,3,1,0} -> levels) ->
         * rational_resampler_fff(5, 1, RRC(5, 24000, 4800, 0.2, 125)) -> x0.66666666 -> frequency_modulator_fc(pi*4800*0.85/24000) ->
         * gr_zero_idle_bursts(delay = (125 - 1) / 2) -> x0.9 -> x bb_gain -> rational_resampler_ccf(sps, 3, low_pass_2(sps, 3 fs, fw,
         * 2000, 60, BH)).  The fft_filter_ccf the constructor also makes (:72-73) is never connected. */
        t->fm = 1; t->sps = 5; t->amplif = 0.9f; t->dmr = 1;
        const float if_samp_rate = 24000, symbol_rate = if_samp_rate / 5.0f;
        int n = qo_firdes_rrc(5, if_samp_rate, symbol_rate, 0.2, 25 * 5, taps, 16384);
        resamp_init(&t->rrc, 1, 5, 1, taps, n);
        t->zi_delay = (unsigned)((n - 1) / 2);
        t->fm_sens = (float)((M_PI * symbol_rate * 0.85) / if_samp_rate);
        /* gr_zero_idle_bursts.cpp:35-38: delay > 0 -> set_history(2 * SAMPLES_PER_SLOT), SAMPLES_PER_SLOT = 720 (src/bursttimer.h:30);
         * the sync block copies in[i] to out[i], and in[0] is the oldest history item: a delay of history - 1 items */
        t->zi_len = t->zi_delay > 0 ? 2 * 720 - 1 : 0;
        t->zi_line = (float*)calloc((size_t)(t->zi_len > 0 ? t->zi_len : 1) * 2, sizeof(float));
        n = qo_firdes_low_pass_2(sps, 3.0 * samp_rate, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->interp, 2, sps, 3, taps, n);
        t->s_sym.isz = 4; qv_init(&t->s_c2, 8);
    } else if (kind == QO_MOD_DSSS) {
        /* /root/reference/src/gr/gr_mod_dsss.cpp:27-93: bits -> scrambler -> cc_encoder -> unpacked_to_packed(1) -> dsss_encoder_bb(barker
         * 13) -> chunks_to_symbols_bc{-1, +1} -> rational_resampler_ccf(sps, 1, RRC(sps, sps, 1, 0.35, 11 sps)) -> x0.65 -> x bb_gain ->
         * rational_resampler_ccf(50, 13, low_pass(50, 5200 * 50, fw, 5 fw)) -> rational_resampler_ccf(50, 1, low_pass(50, fs, fw, 5 fw));
         * the fft_filter_ccf of :67-69 is never connected */
        t->sps = sps; t->amplif = 0.65f; t->dsss = 1;
        int n = qo_firdes_rrc(sps, sps, 1, 0.35, 11 * sps, taps, 16384);
        resamp_init(&t->rrc, 2, sps, 1, taps, n);
        n = qo_firdes_low_pass(50.0, 5200.0 * 50, filter_width, filter_width * 5, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->a_if, 2, 50, 13, taps, n);
        n = qo_firdes_low_pass(50, samp_rate, filter_width, filter_width * 5, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->interp, 2, 50, 1, taps, n);
        t->s_sym.isz = 8; qv_init(&t->s_c2, 8); qv_init(&t->s_clip, 8);
    } else if (kind == QO_MOD_QPSK) {
        /* /root/reference/src/gr/gr_mod_qpsk.cpp:26-90 */
        int nfilts;
        if (sps > 120) nfilts = 11; else if (sps > 10) nfilts = 13; else nfilts = 15;
        t->sps = sps;
        int n = qo_firdes_rrc(sps, sps, 1, 0.35, nfilts * sps, taps, 16384);
        resamp_init(&t->rrc, 2, sps, 1, taps, n);
        t->amplif = 0.6f;
        t->s_sym.isz = 8;
        (void)samp_rate; (void)filter_width;
    } else if (kind == QO_MOD_NBFM) {
        /* /root/reference/src/gr/gr_mod_nbfm.cpp:26-75 (CTCSS tone branch not connected by default) */
        double a[2], b[2];
        qo_preemph_taps(8000, 50e-6, -1.0, a, b);
        iir1_init(&t->preemph, b, a);
        int n = qo_firdes_low_pass_2(1, 8000, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->a_filt, 1, 1, 1, taps, n);
        n = qo_firdes_low_pass_2(25, 50000 * 4, filter_width, 3500, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->a_rs, 1, 25, 4, taps, n);
        t->fm_sens = (float)(4 * M_PI * filter_width / 50000.0f);
        n = qo_firdes_low_pass_2(1, 50000, filter_width, 3500, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->a_if, 2, 1, 1, taps, n);
        n = qo_firdes_low_pass_2(sps, samp_rate, filter_width, 3500, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->interp, 2, sps, 1, taps, n);
        t->amplif = 0.8f;
        t->samp_rate = samp_rate;
        t->a_rs.hkeep = 64; t->a_if.hkeep = 512; t->interp.hkeep = 256;          /* filters set_filter_width may lengthen */
        t->audio_gain = 0.99f; t->a_filt.hkeep = 128;
        qv_init(&t->s_aud, 4); qv_init(&t->s_clip, 8); qv_init(&t->s_c2, 8);
    } else if (kind == QO_MOD_SSB) {
        /* /root/reference/src/gr/gr_mod_ssb.cpp:28-82; flag = sb */
        float tc[2 * 4096];
        int n = qo_firdes_band_pass_2(1, 8000, 300, filter_width, 200, 90, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->a_filt, 1, 1, 1, taps, n);
        n = flag ? qo_firdes_complex_band_pass_2(1, 8000, -filter_width, -200, 200, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096)
                 : qo_firdes_complex_band_pass_2(1, 8000, 200, filter_width, 200, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
        fircc_init(&t->a_sb, tc, n);
        n = qo_firdes_low_pass_2(sps, samp_rate, filter_width, filter_width, 90, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->interp, 2, sps, 1, taps, n);
        t->amplif = 0.9f;
        qv_init(&t->s_aud, 4); qv_init(&t->s_clip, 8); qv_init(&t->s_c2, 8);
        t->samp_rate = samp_rate; t->flag = flag; t->interp.hkeep = 64; t->a_sb.hkeep = 512;
    } else if (kind == QO_MOD_AM) {
        /* /root/reference/src/gr/gr_mod_am.cpp:25-72: audio (8 ksps) -> agc2_ff(1e-2, 1e-4, 1, 1; max gain 1) -> rail_ff(-.98, .98) ->
         * x0.95 -> fft_filter_fff(band_pass_2(1, 8000, 300, 3000, 200, 60, Hamming)) -> + sig_source_f(8000, cos, 0 Hz, 0.5) ->
         * float_to_complex -> rational_resampler_ccf(sps, 1, low_pass(sps, fs, fw, fw)) -> x0.5 -> x bb_gain ->
         * fft_filter_ccc(complex_band_pass_2(1, fs, -fw, fw, 1200, 120, BH)); the feedforward_agc_cc of :51 is never connected.
         * The carrier term: fxpt_nco at phase 0 gives fxpt::cos(0) from the sine table, times the amplitude in double. */
        float tc[2 * 8192];
        agc2_init(&t->am_agc, 1e-2f, 1e-4f, 1.0f, 1.0f); t->am_agc.max_gain = 1.0f;
        int n = qo_firdes_band_pass_2(1, 8000, 300, 3000, 200, 60, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->a_filt, 1, 1, 1, taps, n);
        n = qo_firdes_low_pass(sps, samp_rate, filter_width, filter_width, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->interp, 2, sps, 1, taps, n);
        n = qo_firdes_complex_band_pass_2(1, samp_rate, -filter_width, filter_width, 1200, 120, QO_WIN_BLACKMAN_HARRIS, tc, 8192);
        fircc_init(&t->a_sb, tc, n);
        { const uint32_t uc = 0x40000000u; const int ci = uc >> 22;
          const float c0 = g_sine_tab[2 * ci] * (float)(uc >> 1) + g_sine_tab[2 * ci + 1];
          t->am_dc = (float)((double)c0 * 0.5); }
        t->amplif = 0.5f;
        qv_init(&t->s_aud, 4); qv_init(&t->s_clip, 8); qv_init(&t->s_c2, 8);
        t->samp_rate = samp_rate; t->interp.hkeep = 64; t->a_sb.hkeep = 8192;
    } else if (kind == QO_MOD_BPSK) {
        /* /root/reference/src/gr/gr_mod_bpsk.cpp:27-69 */
        t->sps = sps;
        int n = qo_firdes_rrc(sps, sps, 1, 0.35, 11 * sps, taps, 16384);
        resamp_init(&t->rrc, 2, sps, 1, taps, n);
        t->amplif = 0.6f; t->s_sym.isz = 8;
    } else if (kind == QO_MOD_GMSK) {
        /* /root/reference/src/gr/gr_mod_gmsk.cpp:30-100: the 2FSK (fm) modulator path with a Gaussian pulse (BT 0.3),
         * sensitivity (pi/2)/sps and a x5 (or x1) final interpolation; instances gr_mod_base.cpp:160-162 */
        t->kind = QO_MOD_2FSK; t->fm = 1;
        int nfilts = 35, second_interp = 5;
        if (sps == 10) { sps = 50; second_interp = 1; nfilts = 55; }
        if (sps == 50) nfilts = 55;
        if (sps == 100) nfilts = 35;
        if ((nfilts % 2) == 0) nfilts += 1;
        t->sps = sps; t->amplif = 0.9f;
        int n = qo_firdes_gaussian(sps, sps, 0.3, nfilts, taps, 16384);
        resamp_init(&t->rrc, 1, sps, 1, taps, n);
        t->fm_sens = (float)((M_PI / 2) / sps);
        n = qo_firdes_low_pass(second_interp, samp_rate, filter_width, filter_width, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->interp, 2, second_interp, 1, taps, n);
        t->s_sym.isz = 4;
    } else if (kind == QO_MOD_2FSK) {
        /* /root/reference/src/gr/gr_mod_2fsk.cpp:26-100 */
        int fm = flag; t->fm = fm; t->sps = sps;
        int nfilts = 25 * sps, spacing = 2; t->amplif = 0.8f;
        if (fm) { spacing = 1; t->amplif = 0.9f; }
        if (sps == 5) nfilts = nfilts * 5;
        if ((nfilts % 2) == 0) nfilts += 1;
        int n = qo_firdes_rrc(sps, sps, 1, 0.2, nfilts, taps, 16384);
        resamp_init(&t->rrc, 1, sps, 1, taps, n);
        t->fm_sens = (float)((spacing * M_PI / 2) / sps);
        n = qo_firdes_low_pass(10, samp_rate, filter_width, filter_width, QO_WIN_HAMMING, taps, 16384);
        resamp_init(&t->interp, 2, 10, 1, taps, n);
        t->s_sym.isz = 4;
    } else { free(t); return NULL; }
    return t;
}
void qo_tx_destroy(qo_tx* t)
{
    if (!t) return;
    qv_free(&t->s_bits); qv_free(&t->s_coded); qv_free(&t->s_sym); qv_free(&t->s_shaped); qv_free(&t->s_mod); qv_free(&t->out);
    free(t->zi_line); free(t->zi_tag_off); free(t->zi_tag_val);
    free(t);
}
void qo_tx_set_bb_gain(qo_tx* t, float g) { t->bb_gain = g; }
/* run-time setters of the modulators.  gr_mod_nbfm::set_filter_width (gr_mod_nbfm.cpp:78-93): new taps for the 25/4 resampler
 * (transition width = fw now, not 3500), the 50 ksps filter (transition 1200) and the final interpolator, new modulator
 * sensitivity; like GNU Radio's set_taps the new taps meet the stream's true history from the next output on. */
int qo_tx_set_param(qo_tx* t, int key, double value)
{
    static float taps[16384];
    if (!t) return -1;
    if (key == QO_PARAM_FILTER_WIDTH && t->kind == QO_MOD_NBFM) {
        const int fw = (int)value;
        const float if_samp_rate = 50000;
        int n = qo_firdes_low_pass_2(25, if_samp_rate * 4, fw, fw, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_retap(&t->a_rs, taps, n);
        n = qo_firdes_low_pass_2(1, if_samp_rate, fw, 1200, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_retap(&t->a_if, taps, n);
        n = qo_firdes_low_pass_2(t->interp.L, t->samp_rate, fw, fw, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_retap(&t->interp, taps, n);
        t->fm_sens = (float)(4 * M_PI * fw / if_samp_rate);
        return 0;
    }
    if (key == QO_PARAM_CTCSS && t->kind == QO_MOD_NBFM) {
        /* gr_mod_nbfm::set_ctcss (gr_mod_nbfm.cpp:101-139): 0 -> gain 0.98 (the constructor's is 0.99), low-pass audio filter, tone branch
         * out; f -> gain 0.85, band_pass_2(1, 8000, 300, 3500, 200, 35, BH), the tone source at f added in front of the pre-emphasis.
         * The tone source only runs while it is connected: its phase advances per sample produced with the tone on. */
        int n;
        if (value == 0) {
            t->audio_gain = 0.98f; t->tone_on = 0;
            n = qo_firdes_low_pass_2(1, 8000, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        } else {
            t->audio_gain = 0.85f; t->tone_on = 1;
            n = qo_firdes_band_pass_2(1, 8000, 300, 3500, 200, 35, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
            /* sig_source_f::set_frequency -> fxpt_nco::set_freq((float)(2 pi f / fs)) -> fxpt::float_to_fixed */
            float x = (float)(2 * M_PI * (double)(float)value / 8000.0);
            const int d = (int)floor(x / (float)(2.0 * M_PI) + 0.5);
            x -= d * (float)(2.0 * M_PI);
            t->tone_inc = (uint32_t)(int32_t)((float)x * 2147483648.0f / (float)M_PI);
        }
        resamp_retap(&t->a_filt, taps, n);
        return 0;
    }
    if (key == QO_PARAM_FILTER_WIDTH && t->kind == QO_MOD_SSB) {
        /* gr_mod_ssb::set_filter_width (gr_mod_ssb.cpp:85-100): interpolator low_pass_2(sps, fs, fw, fw, 90) and the side-band filter, now
         * 300 .. fw with a 250 Hz transition (the constructor's is 200 .. fw, 200) */
        static float tc[2 * 4096];
        const int fw = (int)value;
        int n = qo_firdes_low_pass_2(t->interp.L, t->samp_rate, fw, fw, 90, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_retap(&t->interp, taps, n);
        n = t->flag ? qo_firdes_complex_band_pass_2(1, 8000, -fw, -300, 250, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096)
                    : qo_firdes_complex_band_pass_2(1, 8000, 300, fw, 250, 90, QO_WIN_BLACKMAN_HARRIS, tc, 4096);
        fircc_retap(&t->a_sb, tc, n);
        return 0;
    }
    if (key == QO_PARAM_FILTER_WIDTH && t->kind == QO_MOD_AM) {
        /* gr_mod_am::set_filter_width (gr_mod_am.cpp:75-85) */
        static float tc[2 * 8192];
        const int fw = (int)value;
        int n = qo_firdes_low_pass(t->interp.L, t->samp_rate, fw, fw, QO_WIN_HAMMING, taps, 16384);
        resamp_retap(&t->interp, taps, n);
        n = qo_firdes_complex_band_pass_2(1, t->samp_rate, -fw, fw, 1200, 120, QO_WIN_BLACKMAN_HARRIS, tc, 8192);
        fircc_retap(&t->a_sb, tc, n);
        return 0;
    }
    return -1;
}

/* The "zero_samples" stream tag of gr_dmr_source.cpp:148 / gr_mmdvm_source.cpp:264, attached to byte `byte_offset` of the modulator's
 * input with value n_samples.  GNU Radio carries it through packed_to_unpacked (x8), pack_k_bits(2) (/2) and the x5 pulse shaper:
 * at gr_zero_idle_bursts it sits on item 20 * byte_offset.  gr_zero_idle_bursts.cpp:61-70: at output item (offset - delay) the block
 * loads its counter with the value (a later tag overrides a running count; of several tags on one item the first registered wins
 * here, the reference's std::sort leaves that open) and zeroes one output per count.  Deviation, documented: the reference only sees
 * a tag whose item lies at least `delay` items inside the current work() window (its lookup is get_tags_in_window of the CURRENT
 * window matched against item + delay), so it drops tags that fall into the first `delay` items of a scheduler chunk; here a tag is
 * honoured wherever the chunk boundaries are.  A tag registered after its start item has already been produced zeroes what is left
 * of its count.  Returns 0, or -1 when the block has no zero-idle stage. */
int qo_tx_zero_samples(qo_tx* t, long long byte_offset, long n_samples)
{
    if (!t || !t->dmr || byte_offset < 0 || n_samples < 0) return -1;
    long long item = byte_offset * 20;
    if (item < (long long)t->zi_delay) return 0;        /* gr_zero_idle_bursts.cpp:63: offset == nitems + i + delay never holds */
    long long start = item - (long long)t->zi_delay;
    uint64_t val = (uint64_t)n_samples;
    if (start < (long long)t->zi_n) {
        const uint64_t late = (uint64_t)((long long)t->zi_n - start);
        if (late >= val) return 0;
        val -= late; start = (long long)t->zi_n;
    }
    for (long i = 0; i < t->zi_ntags; i++) if (t->zi_tag_off[i] == start) return 0;      /* first registered wins */
    if (t->zi_ntags == t->zi_cap) {
        t->zi_cap = t->zi_cap ? 2 * t->zi_cap : 16;
        t->zi_tag_off = (long long*)realloc(t->zi_tag_off, sizeof(long long) * (size_t)t->zi_cap);
        t->zi_tag_val = (uint64_t*)realloc(t->zi_tag_val, sizeof(uint64_t) * (size_t)t->zi_cap);
    }
    t->zi_tag_off[t->zi_ntags] = start; t->zi_tag_val[t->zi_ntags] = val; t->zi_ntags++;
    return 0;
}
/* gr_zero_idle_bursts::work (gr_zero_idle_bursts.cpp:45-82) on n complex items, in place */
static void zero_idle_work(qo_tx* t, float* x, size_t n);
/* the block alone (tests/test_oracle_ref.py pins it to the reference's compiled gr_zero_idle_bursts.cpp): tags are given on the
 * block's own input items */
void qo_zero_idle_run(const float* in_c, long n, unsigned delay, const long long* tag_item, const long long* tag_val, long ntags, float* out_c)
{
    qo_tx t; memset(&t, 0, sizeof t);
    t.dmr = 1; t.zi_delay = delay; t.zi_len = delay > 0 ? 2 * 720 - 1 : 0;
    t.zi_line = (float*)calloc((size_t)(t.zi_len > 0 ? t.zi_len : 1) * 2, sizeof(float));
    for (long i = 0; i < ntags; i++) {
        if (tag_item[i] < (long long)delay) continue;
        /* same bookkeeping as qo_tx_zero_samples, on block items */
        const long long start = tag_item[i] - (long long)delay;
        int dup = 0;
        for (long k = 0; k < t.zi_ntags; k++) if (t.zi_tag_off[k] == start) dup = 1;
        if (dup) continue;
        if (t.zi_ntags == t.zi_cap) {
            t.zi_cap = t.zi_cap ? 2 * t.zi_cap : 16;
            t.zi_tag_off = (long long*)realloc(t.zi_tag_off, sizeof(long long) * (size_t)t.zi_cap);
            t.zi_tag_val = (uint64_t*)realloc(t.zi_tag_val, sizeof(uint64_t) * (size_t)t.zi_cap);
        }
        t.zi_tag_off[t.zi_ntags] = start; t.zi_tag_val[t.zi_ntags] = (uint64_t)tag_val[i]; t.zi_ntags++;
    }
    memcpy(out_c, in_c, sizeof(float) * 2 * (size_t)n);
    zero_idle_work(&t, out_c, (size_t)n);
    free(t.zi_line); free(t.zi_tag_off); free(t.zi_tag_val);
}
static void zero_idle_work(qo_tx* t, float* x, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        float re = x[2 * i], im = x[2 * i + 1];
        if (t->zi_len > 0) {
            const size_t slot = (size_t)(t->zi_n % (uint64_t)t->zi_len);
            const float dr = t->zi_line[2 * slot], di = t->zi_line[2 * slot + 1];
            t->zi_line[2 * slot] = re; t->zi_line[2 * slot + 1] = im;
            re = dr; im = di;
        }
        for (long k = 0; k < t->zi_ntags; k++)
            if (t->zi_tag_off[k] == (long long)t->zi_n) {
                t->zi_counter = t->zi_tag_val[k];
                t->zi_tag_off[k] = t->zi_tag_off[t->zi_ntags - 1]; t->zi_tag_val[k] = t->zi_tag_val[t->zi_ntags - 1]; t->zi_ntags--;
                break;
            }
        if (t->zi_counter > 0) { re = 0.0f; im = 0.0f; t->zi_counter--; }
        x[2 * i] = re; x[2 * i + 1] = im;
        t->zi_n++;
    }
}

/* analog::frequency_modulator_fc (A13).  Default: Q32 fixed-point phase accumulator (prefix-sum friendly,
 * documented deviation); literal: float accumulator + fmodf as GNU Radio does. */
static void fm_mod(qo_tx* t, const float* x, size_t n, qvec* out, float post)
{
    for (size_t i = 0; i < n; i++) {
        uint32_t ux;
        if (g_fm_literal) {
            const float F_PI = (float)M_PI;
            t->phase_f = t->phase_f + t->fm_sens * x[i];
            t->phase_f = fmodf(t->phase_f + F_PI, 2.0f * F_PI) - F_PI;
            /* gr::fxpt::float_to_fixed */
            float ph = t->phase_f;
            int d = (int)floor(ph / 2 / M_PI + 0.5);
            ph = (float)(ph - d * 2 * M_PI);
            ux = (uint32_t)(int32_t)((float)ph * 2147483648.0f / (float)M_PI);
        } else {
            /* phase increment quantised once to Q32, accumulated exactly modulo 2^32 */
            float inc = t->fm_sens * x[i];
            double q = rint((double)inc * (2147483648.0 / M_PI));
            t->phase_q += (uint32_t)(int32_t)(long long)q;
            ux = t->phase_q;
        }
        /* gr::fxpt::sincos */
        int si = ux >> 22;
        float s = g_sine_tab[2 * si] * (float)(ux >> 1) + g_sine_tab[2 * si + 1];
        uint32_t uc = ux + 0x40000000u;
        int ci = uc >> 22;
        float c = g_sine_tab[2 * ci] * (float)(uc >> 1) + g_sine_tab[2 * ci + 1];
        qv_pushc(out, c * post, s * post);
    }
}

int qo_tx_work(qo_tx* t, const void* in, long n)
{
    const uint8_t* bytes = (const uint8_t*)in;
    if (t->kind == QO_MOD_DSSS) {
        static const int barker_13[13] = { 1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 0, 1 };
        t->s_bits.n = 0;
        for (long i = 0; i < n; i++)
            for (int b = 7; b >= 0; b--) qv_pushb(&t->s_bits, lfsr_scramble(&t->scr, (bytes[i] >> b) & 1));
        t->s_coded.n = 0;
        ccenc_work(&t->enc, t->s_bits.d, t->s_bits.n, &t->s_coded);
        /* unpacked_to_packed (8 coded bits per byte, MSB first) then dsss_encoder_bb walks the byte MSB first again: per coded bit
         * 13 chips, the code for a 0 and its complement for a 1 (dsss_encoder_bb_impl.cc:86-97); chips -> {-1, +1} */
        t->s_sym.n = 0;
        for (size_t i = 0; i < t->s_coded.n; i++)
            for (int c = 0; c < 13; c++) {
                const int chip = t->s_coded.d[i] == 0 ? (1 & barker_13[c]) : (1 & ~barker_13[c]);
                qv_pushc(&t->s_sym, chip ? 1.0f : -1.0f, 0.0f);
            }
        t->s_c2.n = 0; resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->s_c2);
        float* m = (float*)t->s_c2.d;
        for (size_t i = 0; i < 2 * t->s_c2.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        t->s_clip.n = 0; resamp_work(&t->a_if, m, t->s_c2.n, &t->s_clip);
        resamp_work(&t->interp, (const float*)t->s_clip.d, t->s_clip.n, &t->out);
        return 0;
    }
    if (t->kind == QO_MOD_M17 || t->kind == QO_MOD_DMR) {
        static const int map[4] = { 2, 3, 1, 0 };
        static const float lv[4] = { -1.5f, -0.5f, 0.5f, 1.5f };
        t->s_sym.n = 0;
        for (long i = 0; i < n; i++)
            for (int b = 6; b >= 0; b -= 2) qv_pushf(&t->s_sym, lv[map[(bytes[i] >> b) & 3]]);
        t->s_shaped.n = 0;
        resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->s_shaped);
        float* p = (float*)t->s_shaped.d;
        for (size_t i = 0; i < t->s_shaped.n; i++) p[i] = p[i] * 0.66666666f;
        t->s_mod.n = 0;
        fm_mod(t, p, t->s_shaped.n, &t->s_mod, 1.0f);
        t->s_c2.n = 0;
        if (t->dmr) { zero_idle_work(t, (float*)t->s_mod.d, t->s_mod.n); qv_push(&t->s_c2, t->s_mod.d, t->s_mod.n); }
        else resamp_work(&t->a_if, (const float*)t->s_mod.d, t->s_mod.n, &t->s_c2);
        float* m = (float*)t->s_c2.d;
        for (size_t i = 0; i < 2 * t->s_c2.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        resamp_work(&t->interp, m, t->s_c2.n, &t->out);
        return 0;
    }
    if (t->kind == QO_MOD_4FSK || t->kind == QO_MOD_QPSK) {
        /* packed_to_unpacked(1, MSB) -> scrambler -> cc_encoder */
        t->s_bits.n = 0;
        for (long i = 0; i < n; i++)
            for (int b = 7; b >= 0; b--) qv_pushb(&t->s_bits, lfsr_scramble(&t->scr, (bytes[i] >> b) & 1));
        t->s_coded.n = 0;
        ccenc_work(&t->enc, t->s_bits.d, t->s_bits.n, &t->s_coded);
        /* pack_k_bits(2) MSB first -> map {0,1,3,2} -> symbols */
        static const int map[4] = { 0, 1, 3, 2 };
        t->s_sym.n = 0;
        for (size_t i = 0; i + 1 < t->s_coded.n; i += 2) {
            int chunk = map[(t->s_coded.d[i] << 1) | t->s_coded.d[i + 1]];
            if (t->kind == QO_MOD_4FSK) {
                static const float lv[4] = { -1.5f, -0.5f, 0.5f, 1.5f };
                qv_pushf(&t->s_sym, lv[chunk]);
            } else {
                /* digital::diff_encoder_bb(4) then chunks_to_symbols_bc */
                t->diff_prev = (chunk + t->diff_prev) % 4;
                static const float qr[4] = { -0.707f, -0.707f, 0.707f, 0.707f };
                static const float qi[4] = { -0.707f, 0.707f, 0.707f, -0.707f };
                qv_pushc(&t->s_sym, qr[t->diff_prev], qi[t->diff_prev]);
            }
        }
        if (t->kind == QO_MOD_4FSK) {
            t->s_shaped.n = 0;
            if (t->fm) {
                resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->s_shaped);
                float* p = (float*)t->s_shaped.d;
                for (size_t i = 0; i < t->s_shaped.n; i++) p[i] = p[i] * 0.66666666f;
            } else {
                const float* p = (const float*)t->s_sym.d;
                for (size_t i = 0; i < t->s_sym.n; i++) for (int k = 0; k < t->sps; k++) qv_pushf(&t->s_shaped, p[i]);
            }
            t->s_mod.n = 0;
            fm_mod(t, (const float*)t->s_shaped.d, t->s_shaped.n, &t->s_mod, 1.0f);
            float* m = (float*)t->s_mod.d;
            for (size_t i = 0; i < 2 * t->s_mod.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
            resamp_work(&t->interp, m, t->s_mod.n, &t->out);
        } else {
            size_t o0 = t->out.n;
            resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->out);
            float* m = (float*)t->out.d;
            for (size_t i = 2 * o0; i < 2 * t->out.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        }
        return 0;
    }
    if (t->kind == QO_MOD_NBFM) {
        /* n float audio samples at 8 ksps */
        const float* au = (const float*)in;
        t->s_aud.n = 0; resamp_work(&t->a_filt, au, (size_t)n, &t->s_aud);
        float* a1 = (float*)t->s_aud.d;
        for (size_t i = 0; i < t->s_aud.n; i++) {
            a1[i] = a1[i] * t->audio_gain;                                                   /* multiply_const_ff(0.99 / 0.98 / 0.85) */
            if (t->tone_on) {
                /* add_ff with sig_source_f(8000, cos, f, 0.15): fxpt_nco: (float)(fxpt::cos(phase) * ampl), then phase += inc */
                const uint32_t uc = t->tone_phase + 0x40000000u;
                const int ci = uc >> 22;
                const float cs = g_sine_tab[2 * ci] * (float)(uc >> 1) + g_sine_tab[2 * ci + 1];
                a1[i] = a1[i] + (float)((double)cs * 0.15);
                t->tone_phase += t->tone_inc;
            }
        }
        t->s_shaped.n = 0; iir1_work(&t->preemph, a1, t->s_aud.n, &t->s_shaped, 1.0f);
        t->s_sym.isz = 4; t->s_sym.n = 0;
        resamp_work(&t->a_rs, (const float*)t->s_shaped.d, t->s_shaped.n, &t->s_sym);       /* 8k -> 50k */
        t->s_mod.n = 0; fm_mod(t, (const float*)t->s_sym.d, t->s_sym.n, &t->s_mod, 1.0f);
        t->s_c2.n = 0; resamp_work(&t->a_if, (const float*)t->s_mod.d, t->s_mod.n, &t->s_c2);
        float* m = (float*)t->s_c2.d;
        for (size_t i = 0; i < 2 * t->s_c2.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        resamp_work(&t->interp, m, t->s_c2.n, &t->out);
        return 0;
    }
    if (t->kind == QO_MOD_AM) {
        const float* au = (const float*)in;
        t->s_aud.n = 0;
        for (long i = 0; i < n; i++) {
            /* analog::kernel::agc2_ff::scale (fabsf form), then rail_ff, then multiply_const_ff(0.95) */
            const float out = au[i] * t->am_agc.gain;
            const float tmp = fabsf(out) - t->am_agc.ref;
            float rate = t->am_agc.decay;
            if (fabsf(tmp) > t->am_agc.gain) rate = t->am_agc.attack;
            t->am_agc.gain -= tmp * rate;
            if (t->am_agc.gain < 0.0f) t->am_agc.gain = 10e-5f;
            if (t->am_agc.max_gain > 0.0f && t->am_agc.gain > t->am_agc.max_gain) t->am_agc.gain = t->am_agc.max_gain;
            float v = out < -0.98f ? -0.98f : (out > 0.98f ? 0.98f : out);
            v = v * 0.95f;
            qv_pushf(&t->s_aud, v);
        }
        t->s_shaped.n = 0; resamp_work(&t->a_filt, (const float*)t->s_aud.d, t->s_aud.n, &t->s_shaped);
        const float* f = (const float*)t->s_shaped.d;
        t->s_clip.n = 0;
        for (size_t i = 0; i < t->s_shaped.n; i++) qv_pushc(&t->s_clip, f[i] + t->am_dc, 0.0f);       /* add_ff, float_to_complex */
        t->s_c2.n = 0; resamp_work(&t->interp, (const float*)t->s_clip.d, t->s_clip.n, &t->s_c2);
        float* m = (float*)t->s_c2.d;
        for (size_t i = 0; i < 2 * t->s_c2.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        fircc_work(&t->a_sb, m, t->s_c2.n, &t->out);
        return 0;
    }
    if (t->kind == QO_MOD_SSB) {
        const float* au = (const float*)in;
        t->s_aud.n = 0; resamp_work(&t->a_filt, au, (size_t)n, &t->s_aud);
        const float* a1 = (const float*)t->s_aud.d;
        for (size_t i = 0; i < t->s_aud.n; i++) {                                           /* float_to_complex -> clipper_cc(0.95) */
            float cr, ci;
            cessb_clip_one(a1[i], 0.0f, 0.95f, &cr, &ci);
            qv_pushc(&t->s_clip, cr, ci);
        }
        const float* c = (const float*)t->s_clip.d;
        t->s_c2.n = 0;
        while (t->st_pos + 2 < t->s_clip.n) {                                                /* stretcher_cc */
            size_t k = t->st_pos;
            float e0 = sqrtf(c[2 * k] * c[2 * k] + c[2 * k + 1] * c[2 * k + 1]);
            float e1 = sqrtf(c[2 * (k + 1)] * c[2 * (k + 1)] + c[2 * (k + 1) + 1] * c[2 * (k + 1) + 1]);
            float e2 = sqrtf(c[2 * (k + 2)] * c[2 * (k + 2)] + c[2 * (k + 2) + 1] * c[2 * (k + 2) + 1]);
            float h = cessb_stretch_div(t->env_m2, t->env_m1, e0, e1, e2);
            qv_pushc(&t->s_c2, c[2 * k] / h, c[2 * k + 1] / h);
            t->env_m2 = t->env_m1; t->env_m1 = e0;
            t->st_pos++;
        }
        if (t->st_pos > 0) { qv_drop(&t->s_clip, t->st_pos); t->st_pos = 0; }
        t->s_mod.n = 0; fircc_work(&t->a_sb, (const float*)t->s_c2.d, t->s_c2.n, &t->s_mod);
        float* m = (float*)t->s_mod.d;
        for (size_t i = 0; i < 2 * t->s_mod.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        resamp_work(&t->interp, m, t->s_mod.n, &t->out);
        return 0;
    }
    if (t->kind == QO_MOD_BPSK || t->kind == QO_MOD_2FSK) {
        t->s_bits.n = 0;
        for (long i = 0; i < n; i++)
            for (int b = 7; b >= 0; b--) qv_pushb(&t->s_bits, lfsr_scramble(&t->scr, (bytes[i] >> b) & 1));
        t->s_coded.n = 0;
        ccenc_work(&t->enc, t->s_bits.d, t->s_bits.n, &t->s_coded);
        t->s_sym.n = 0;
        for (size_t i = 0; i < t->s_coded.n; i++) {
            float lv = t->s_coded.d[i] ? 1.0f : -1.0f;
            if (t->kind == QO_MOD_BPSK) qv_pushc(&t->s_sym, lv, 0.0f); else qv_pushf(&t->s_sym, lv);
        }
        if (t->kind == QO_MOD_BPSK) {
            size_t o0 = t->out.n;
            resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->out);
            float* m = (float*)t->out.d;
            for (size_t i = 2 * o0; i < 2 * t->out.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
        } else {
            t->s_shaped.n = 0;
            if (t->fm) resamp_work(&t->rrc, (const float*)t->s_sym.d, t->s_sym.n, &t->s_shaped);
            else { const float* p = (const float*)t->s_sym.d; for (size_t i = 0; i < t->s_sym.n; i++) for (int k = 0; k < t->sps; k++) qv_pushf(&t->s_shaped, p[i]); }
            t->s_mod.n = 0;
            fm_mod(t, (const float*)t->s_shaped.d, t->s_shaped.n, &t->s_mod, 1.0f);
            float* m = (float*)t->s_mod.d;
            for (size_t i = 0; i < 2 * t->s_mod.n; i++) { m[i] = m[i] * t->amplif; m[i] = m[i] * t->bb_gain; }
            resamp_work(&t->interp, m, t->s_mod.n, &t->out);
        }
        return 0;
    }
    return -1;
}
long qo_tx_out_items(qo_tx* t) { return (long)t->out.n; }
const float* qo_tx_out_data(qo_tx* t) { return (const float*)t->out.d; }
void qo_tx_out_clear(qo_tx* t) { t->out.n = 0; }

/* ------------------------------------------------------------------ framing (gr_modem.cpp:1119-1282 restated) */
/* bit-serial shift-register sync search followed by frame_len bytes packed MSB first (packBytes, gr_modem.cpp:980-994) */
long qo_find_frames(const uint8_t* bits, long nbits, uint32_t sync, int sync_bits, int frame_len, uint8_t* frames, long max_frames)
{
    uint64_t sh = 0; uint64_t mask = (sync_bits >= 64) ? ~0ull : ((1ull << sync_bits) - 1);
    long found = 0; long i = 0;
    while (i < nbits && found < max_frames) {
        sh = ((sh << 1) | (bits[i] & 1)) & mask; i++;
        if (sh == (uint64_t)sync) {
            if (i + (long)frame_len * 8 > nbits) break;
            for (int b = 0; b < frame_len; b++) {
                int t = 0;
                for (int k = 0; k < 8; k++) t = (t << 1) | (bits[i + b * 8 + k] & 1);
                frames[found * frame_len + b] = (uint8_t)t;
            }
            found++; i += (long)frame_len * 8; sh = 0;
        }
    }
    return found;
}

/* ------------------------------------------------------------------ polyphase channelizer / synthesizer
 * gr::filter::pfb_channelizer_ccf(M, taps, 1.0) fed by blocks::stream_to_streams(M)
 *   (/root/reference/src/gr/gr_demod_mmdvm_multi2.cpp:98-107) and gr::filter::pfb_synthesizer_ccf(M, taps, false)
 *   (/root/reference/src/gr/gr_mod_mmdvm_multi2.cpp:90-92), restated from GNU Radio 3.10 (gr-filter
 *   pfb_channelizer_ccf_impl::general_work, polyphase_filterbank::set_taps, pfb_synthesizer_ccf_impl::work; Appendix A):
 *     branch filter k holds taps[k + t*M] (prototype zero-padded to a multiple of M);
 *     channelizer: stream j = x[m*M + j]; u_k[m] = sum_t taps[k + tM] * x[(m - t)M + (M-1-k)];
 *                  out_c[m] = sum_k u_k[m] * exp(+j 2 pi k c / M)        (FFTW backward, unnormalised)
 *     synthesizer: v_i[n] = sum_c in_c[n] * exp(+j 2 pi i c / M);  y[nM + i] = sum_t taps[i + tM] * v_i[n - t]
 *                  (polyphase interpolator: channel c leaves at +c fs/M with zero phase offset, so that
 *                  channelizer(synthesizer(z)) returns channel c on port c -- the convention both reference wirings
 *                  rely on, gr_mod_mmdvm_multi2.cpp:107-121 / gr_demod_mmdvm_multi2.cpp:110-124; upstream's internal
 *                  bin/commutator ordering cannot be re-read offline and is not claimed)
 *   Arithmetic order adopted by this oracle (FFTW's butterfly order is not reproducible offline: parity unpinned, the
 *   float tolerance of 1e-5 RMS applies): branch FIRs accumulate oldest sample first with fmaf; the M-point DFT is the
 *   direct sum over k ascending with twiddles (float)cos / (float)sin of the double angle 2 pi ((k c) mod M) / M and
 *   re = fmaf(ur, wr, re); re = fmaf(-ui, wi, re); im = fmaf(ur, wi, im); im = fmaf(ui, wr, im). */
struct qo_pfb {
    int M, tpf, synth;
    float* bt;              /* [M][tpf] branch taps */
    float* w;               /* [M][2] twiddles */
    qvec in;                /* channelizer: pending input samples incl. (tpf-1)*M history; synthesizer: per-branch v history */
    float* vh;              /* synthesizer: [M][tpf-1][2] history of v_i, oldest first */
};
typedef struct qo_pfb qo_pfb;

static qo_pfb* pfb_create(int M, const float* taps, int ntaps, int synth)
{
    qo_pfb* p = (qo_pfb*)calloc(1, sizeof(qo_pfb));
    p->M = M; p->synth = synth;
    p->tpf = (ntaps + M - 1) / M;
    p->bt = (float*)calloc((size_t)M * p->tpf, sizeof(float));
    for (int j = 0; j < ntaps; j++) p->bt[(size_t)(j % M) * p->tpf + j / M] = taps[j];
    p->w = (float*)calloc((size_t)M * 2, sizeof(float));
    for (int q = 0; q < M; q++) {
        const double a = 2.0 * M_PI * (double)q / (double)M;
        p->w[2 * q] = (float)cos(a); p->w[2 * q + 1] = (float)sin(a);
    }
    qv_init(&p->in, 8);
    if (!synth) qv_push_zero(&p->in, (size_t)(p->tpf - 1) * M);
    else p->vh = (float*)calloc((size_t)M * (p->tpf > 1 ? p->tpf - 1 : 1) * 2, sizeof(float));
    return p;
}
qo_pfb* qo_pfb_channelizer_create(int M, const float* taps, int ntaps) { return pfb_create(M, taps, ntaps, 0); }
qo_pfb* qo_pfb_synthesizer_create(int M, const float* taps, int ntaps) { return pfb_create(M, taps, ntaps, 1); }
void qo_pfb_destroy(qo_pfb* p) { if (!p) return; free(p->bt); free(p->w); free(p->vh); qv_free(&p->in); free(p); }

static void pfb_dft(const qo_pfb* p, const float* u /* [M][2] */, int c, float* re_out, float* im_out)
{
    float re = 0.0f, im = 0.0f;
    for (int k = 0; k < p->M; k++) {
        const int q = (int)(((long)k * c) % p->M);
        const float wr = p->w[2 * q], wi = p->w[2 * q + 1], ur = u[2 * k], ui = u[2 * k + 1];
        re = fmaf(ur, wr, re); re = fmaf(-ui, wi, re);
        im = fmaf(ur, wi, im); im = fmaf(ui, wr, im);
    }
    *re_out = re; *im_out = im;
}

/* x: n complex samples of the wideband stream; out: [M][cap] complex, appended from column `have`; returns the number
 * of new columns (output samples per channel).  Samples that do not fill a frame of M wait for the next call. */
long qo_pfb_channelizer_work(qo_pfb* p, const float* x, long n, float* out, long cap, long have)
{
    const int M = p->M, tpf = p->tpf;
    qv_push(&p->in, x, (size_t)n);
    const float* s = (const float*)p->in.d;                 /* s[0] = oldest history sample */
    const long hist = (long)(tpf - 1) * M;
    const long frames = ((long)p->in.n - hist) / M;
    float* u = (float*)malloc(sizeof(float) * 2 * M);
    for (long m = 0; m < frames && have + m < cap; m++) {
        for (int k = 0; k < M; k++) {
            float re = 0.0f, im = 0.0f;
            for (int t = tpf - 1; t >= 0; t--) {            /* oldest sample first */
                const long idx = hist + (m - t) * M + (M - 1 - k);
                const float h = p->bt[(size_t)k * tpf + t];
                re = fmaf(h, s[2 * idx], re); im = fmaf(h, s[2 * idx + 1], im);
            }
            u[2 * k] = re; u[2 * k + 1] = im;
        }
        for (int c = 0; c < M; c++) pfb_dft(p, u, c, &out[2 * ((size_t)c * cap + have + m)], &out[2 * ((size_t)c * cap + have + m) + 1]);
    }
    free(u);
    qv_drop(&p->in, (size_t)frames * M);
    return frames;
}

/* in: [M][stride] complex (n columns used); out: n*M complex samples of the wideband stream */
long qo_pfb_synthesizer_work(qo_pfb* p, const float* in, long n, long stride, float* out)
{
    const int M = p->M, tpf = p->tpf, H = tpf - 1;
    float* xin = (float*)malloc(sizeof(float) * 2 * M);
    for (long t0 = 0; t0 < n; t0++) {
        for (int c = 0; c < M; c++) { xin[2 * c] = in[2 * ((size_t)c * stride + t0)]; xin[2 * c + 1] = in[2 * ((size_t)c * stride + t0) + 1]; }
        for (int i = 0; i < M; i++) {
            float vr, vi;
            pfb_dft(p, xin, i, &vr, &vi);
            float* hist = p->vh + (size_t)i * (H > 0 ? H : 1) * 2;        /* hist[0] = v_i[n - H] ... hist[H-1] = v_i[n - 1] */
            float re = 0.0f, im = 0.0f;
            for (int t = tpf - 1; t >= 1; t--) {            /* oldest first */
                const float h = p->bt[(size_t)i * tpf + t];
                re = fmaf(h, hist[2 * (H - t)], re); im = fmaf(h, hist[2 * (H - t) + 1], im);
            }
            { const float h = p->bt[(size_t)i * tpf]; re = fmaf(h, vr, re); im = fmaf(h, vi, im); }
            out[2 * ((size_t)t0 * M + i)] = re; out[2 * ((size_t)t0 * M + i) + 1] = im;
            if (H > 0) { memmove(hist, hist + 2, sizeof(float) * 2 * (H - 1)); hist[2 * (H - 1)] = vr; hist[2 * (H - 1) + 1] = vi; }
        }
    }
    free(xin);
    return n * M;
}

/* ------------------------------------------------------------------ layer-1 deframer
 * gr_modem::synchronize + findSync + packBytes (/root/reference/src/gr_modem.cpp:1119-1282, 980-994) restated bit for
 * bit: a shift register searches the sync words of layer1framing.h:8-24; once one is found the next bit_len bits are
 * collected, packed MSB first and handed on with the frame type; then the shift register is cleared.
 * sync_class selects the findSync branch: 1 = "1K" modes (8-bit 0xB5 only; gr_modem.cpp:1208-1221),
 * 2 = narrow modes (16-bit voice 0xED89, 24-bit text / proto / video / callsign / end; :1222-1257),
 * 3 = wide modes QPSK250K / QPSKVideo / 4FSK100K (24-bit IP / video / end; :1258-1274),
 * 4 = M17 (16-bit link-setup 0x55F7 / stream 0xFF5D, 32-bit end of transmission 0x555D555D; :1187-1207).
 * Lengths follow synchronize(): for classes 2, 3 a voice frame takes bit_buf_len bits into rx_frame_length + 1 bytes,
 * any other frame bit_buf_len - 8 bits into rx_frame_length bytes; class 1 always bit_buf_len bits (:1146-1167). */
enum { QO_FT_NONE = 0, QO_FT_VOICE = 0xED89, QO_FT_VOICE1 = 0xB5, QO_FT_TEXT = 0x89EDAA, QO_FT_IP = 0xDE98AA, QO_FT_VIDEO = 0x98DEAA,
       QO_FT_CALLSIGN = 0x8CC8DD, QO_FT_PROTO = 0xED77AA, QO_FT_END = 0x4C8A2B,
       QO_FT_M17_STREAM = 0xFF5D, QO_FT_M17_LSF = 0x55F7, QO_FT_M17_EOT = 0x555D555D };       /* layer1framing.h:21-23 */
struct qo_deframer {
    int sync_class, bit_buf_len, rx_frame_length;
    uint64_t shift_reg; int sync_found; uint32_t cur_type; int bit_idx; int modem_sync;
    uint8_t* bit_buf;
};
typedef struct qo_deframer qo_deframer;

qo_deframer* qo_deframer_create(int sync_class, int bit_buf_len, int rx_frame_length)
{
    qo_deframer* d = (qo_deframer*)calloc(1, sizeof(qo_deframer));
    d->sync_class = sync_class; d->bit_buf_len = bit_buf_len; d->rx_frame_length = rx_frame_length;
    d->bit_buf = (uint8_t*)calloc((size_t)bit_buf_len + 8, 1);
    return d;
}
void qo_deframer_destroy(qo_deframer* d) { if (d) { free(d->bit_buf); free(d); } }
int qo_deframer_modem_sync(const qo_deframer* d) { return d->modem_sync; }

static uint32_t deframer_find_sync(qo_deframer* d, unsigned bit)
{
    d->shift_reg = (d->shift_reg << 1) | (bit & 1u);
    uint64_t t;
    if (d->sync_class == 1) {
        t = d->shift_reg & 0xFF;
        if (t == QO_FT_VOICE1) { d->sync_found = 1; return QO_FT_VOICE1; }
        return QO_FT_NONE;
    }
    if (d->sync_class == 4) {
        /* ModemTypeM17 (gr_modem.cpp:1187-1207): 16-bit link-setup / stream words, else the 32-bit end-of-transmission word */
        t = d->shift_reg & 0xFFFF;
        if (t == QO_FT_M17_LSF) { d->sync_found = 1; return QO_FT_M17_LSF; }
        if (t == QO_FT_M17_STREAM) { d->sync_found = 1; return QO_FT_M17_STREAM; }
        t = d->shift_reg & 0xFFFFFFFFull;
        if (t == QO_FT_M17_EOT) { d->sync_found = 1; return QO_FT_M17_EOT; }
        return QO_FT_NONE;
    }
    if (d->sync_class == 2) {
        t = d->shift_reg & 0xFFFF;
        if (t == QO_FT_VOICE) { d->sync_found = 1; return QO_FT_VOICE; }
        t = d->shift_reg & 0xFFFFFF;
        if (t == QO_FT_TEXT || t == QO_FT_PROTO || t == QO_FT_VIDEO || t == QO_FT_CALLSIGN || t == QO_FT_END) { d->sync_found = 1; return (uint32_t)t; }
        return QO_FT_NONE;
    }
    t = d->shift_reg & 0xFFFFFF;
    if (t == QO_FT_IP || t == QO_FT_VIDEO || t == QO_FT_END) { d->sync_found = 1; return (uint32_t)t; }
    return QO_FT_NONE;
}

/* bits: one bit per byte.  records: max_frames records of rec_bytes each = { u32 type, u32 nbytes, payload... };
 * returns the number of frames completed by this call (state carries over). */
long qo_deframer_work(qo_deframer* d, const uint8_t* bits, long n, uint8_t* records, int rec_bytes, long max_frames)
{
    long found = 0;
    for (long i = 0; i < n; i++) {
        if (!d->sync_found) {
            d->cur_type = deframer_find_sync(d, bits[i]);
            if (d->sync_found) { d->bit_idx = 0; if (d->modem_sync < 32) d->modem_sync += 8; continue; }
            else if (d->modem_sync > 0) d->modem_sync -= 1;
        }
        if (d->sync_found) {
            d->bit_buf[d->bit_idx++] = bits[i] & 1;
            int frame_length = d->rx_frame_length, bit_len = d->bit_buf_len;
            if (d->sync_class != 1 && d->sync_class != 4) {      /* the "1K" modes and M17 always take bit_buf_len bits (gr_modem.cpp:1146-1167) */
                if (d->cur_type == QO_FT_VOICE) frame_length++;     /* reserved byte */
                else bit_len = d->bit_buf_len - 8;
            }
            if (d->bit_idx >= bit_len) {
                if (found < max_frames) {
                    uint8_t* r = records + (size_t)found * rec_bytes;
                    memset(r, 0, (size_t)rec_bytes);
                    const uint32_t ty = d->cur_type, nb = (uint32_t)frame_length;
                    memcpy(r, &ty, 4); memcpy(r + 4, &nb, 4);
                    for (int b = 0; b * 8 < bit_len && 8 + b < rec_bytes; b++) {
                        int t = 0;
                        for (int k = 0; k < 8; k++) t = (t << 1) | (d->bit_buf[b * 8 + k] & 1);
                        r[8 + b] = (uint8_t)t;
                    }
                    found++;
                }
                d->sync_found = 0; d->shift_reg = 0; d->bit_idx = 0;
            }
        }
    }
    return found;
}

/* ------------------------------------------------------------------ gr_deframer_bb (MMDVM-era bit deframer behind ports 2 / 3 of
 * the dual-decoder modes: gr_demod_base.cpp wires _deframer1/2 (type 1), _deframer_700_1/2 (type 2), _deframer_10k_1/2 (type 3))
 * /root/reference/src/gr/gr_deframer_bb.cpp:83-185 restated bit for bit, quirks included: the 8- or 16-bit word under test is
 * compared against every 16-bit sync word (type 2 masks 8 bits, so only 0xB5 and the 24-bit End word can hit there); on a hit the
 * block emits the matched word MSB first -- 16 bits, 24 for the End word, 8 for type 2 (the LOW 8 bits of whatever matched) --
 * then the next bit_buf_len input bits verbatim, then clears its shift register.  Output: a bit stream (one bit per byte). */
struct qo_dfbb { int type, sync_found, idx, len; uint64_t shift; };
typedef struct qo_dfbb qo_dfbb;
qo_dfbb* qo_dfbb_create(int modem_type)
{
    if (modem_type < 1 || modem_type > 3) return NULL;          /* the reference leaves _bit_buf_len uninitialised otherwise */
    qo_dfbb* d = (qo_dfbb*)calloc(1, sizeof(qo_dfbb));
    d->type = modem_type;
    d->len = modem_type == 1 ? 8 * 8 : (modem_type == 2 ? 4 * 8 : 48 * 8);
    return d;
}
void qo_dfbb_destroy(qo_dfbb* d) { free(d); }
static uint32_t dfbb_find_sync(qo_dfbb* d, unsigned bit)
{
    d->shift = (d->shift << 1) | (bit & 1u);
    uint32_t t = (uint32_t)(d->type != 2 ? (d->shift & 0xFFFF) : (d->shift & 0xFF));
    if (d->type == 2 && t == 0xB5) { d->sync_found = 1; return t; }
    if (t == 0x89ED || t == 0xED89 || t == 0x98DE || t == 0xED77 || t == 0x8CC8) { d->sync_found = 1; return t; }
    t = (uint32_t)(d->shift & 0xFFFFFF);
    if (t == 0x4C8A2B) { d->sync_found = 1; return t; }
    return 0;
}
/* returns the number of output bits appended to out (at most cap; the rest is dropped, state still advances) */
long qo_dfbb_work(qo_dfbb* d, const uint8_t* bits, long n, uint8_t* out, long cap)
{
    long no = 0;
    for (long i = 0; i < n; i++) {
        if (!d->sync_found) {
            const uint32_t ft = dfbb_find_sync(d, bits[i]);
            if (d->sync_found) {
                int nb;
                if ((d->type == 1 || d->type == 3) && ft != 0x4C8A2B) nb = 16;
                else if ((d->type == 1 || d->type == 3) && ft == 0x4C8A2B) nb = 24;
                else nb = 8;
                for (int k = 0; k < nb; k++) { if (no < cap) out[no] = (uint8_t)((ft >> (nb - 1 - k)) & 1u); no++; }
                d->idx = 0;
                continue;
            }
        }
        if (d->sync_found) {
            if (no < cap) out[no] = bits[i] & 1u;
            no++;
            if (++d->idx >= d->len) { d->sync_found = 0; d->shift = 0; d->idx = 0; }
        }
    }
    return no < cap ? no : cap;
}

/* ------------------------------------------------------------------ gr_modem::frame (TX framing)
 * /root/reference/src/gr_modem.cpp:904-961: [10 x 0xAA when an IP frame goes out in burst mode] + the frame type's sync word
 * (voice: 0xB5 for the "1K" modes, else 0xED89 + the reserved byte 0xAA; text / video / IP / proto: their 24-bit words; any other
 * type: nothing) + the payload.  Returns the number of bytes written (at most cap). */
long qo_frame(const uint8_t* payload, long n, uint32_t frame_type, int one_k_mode, int burst_ip, uint8_t* out, long cap)
{
    long k = 0;
#define QO_PUT(v) do { if (k < cap) out[k] = (uint8_t)(v); k++; } while (0)
    if (frame_type == QO_FT_IP && burst_ip) for (int i = 0; i < 10; i++) QO_PUT(0xAA);
    if (frame_type == QO_FT_VOICE) {
        if (one_k_mode) QO_PUT(QO_FT_VOICE1 & 0xFF);
        else { QO_PUT((QO_FT_VOICE >> 8) & 0xFF); QO_PUT(QO_FT_VOICE & 0xFF); QO_PUT(0xAA); }
    } else if (frame_type == QO_FT_TEXT || frame_type == QO_FT_VIDEO || frame_type == QO_FT_IP || frame_type == QO_FT_PROTO) {
        QO_PUT((frame_type >> 16) & 0xFF); QO_PUT((frame_type >> 8) & 0xFF); QO_PUT(frame_type & 0xFF);
    }
    for (long i = 0; i < n; i++) QO_PUT(payload[i]);
#undef QO_PUT
    return k < cap ? k : cap;
}

/* ------------------------------------------------------------------ RSSI tap
 * /root/reference/src/gr/rssi_block.cpp:25-45 on the demodulators' port 0 (gr_demod_base.cpp:199-200):
 * complex_to_mag_squared -> moving_average_ff(2000, 1, 2000) -> single_pole_iir_filter_ff(0.04) -> 10 log10 -> + level,
 * read through probe_signal_f (the latest value).  moving_average_ff: sum += newest; out = sum * scale; sum -= the
 * sample length-1 older (float running sum, zero history); single_pole_iir<float,float,double>:
 * y = alpha x + (1 - alpha) y_prev in double, rounded to float.  nlog10_ff clamps its argument at 1e-18; this oracle
 * uses log10f (VOLK's polynomial log2 is not restated: telemetry value, tolerance 1e-3 dB in the tests). */
struct qo_rssi { float ring[2000]; int head; float sum; double y_prev; float last_db; float level; long n; };
typedef struct qo_rssi qo_rssi;
qo_rssi* qo_rssi_create(float level) { qo_rssi* r = (qo_rssi*)calloc(1, sizeof(qo_rssi)); r->level = level; r->last_db = 10.0f * log10f(1e-18f) + level; return r; }
void qo_rssi_destroy(qo_rssi* r) { free(r); }
float qo_rssi_work(qo_rssi* r, const float* iq, long n)
{
    for (long i = 0; i < n; i++) {
        const float m2 = iq[2 * i] * iq[2 * i] + iq[2 * i + 1] * iq[2 * i + 1];
        /* ring[head] holds the sample 1999 older than the newest once the history is full (zeros before) */
        r->sum = r->sum + m2;
        const float out = r->sum * 1.0f;
        r->ring[(r->head + 1999) % 2000] = m2;           /* slot of the newest */
        r->sum = r->sum - r->ring[r->head];               /* the sample length-1 = 1999 older than the newest */
        r->head = (r->head + 1) % 2000;
        const double y = 0.04 * (double)out + (1.0 - 0.04) * r->y_prev;
        const float yf = (float)y;
        r->y_prev = (double)yf;
        r->last_db = 10.0f * log10f(yf > 1e-18f ? yf : 1e-18f) + r->level;
        r->n++;
    }
    return r->last_db;
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * rx_fft_c (/root/reference/src/gr/rx_fft.cpp:44-129), the display spectrum: samples x window fill an N-item buffer; when the
 * buffer is full at the NEXT incoming sample the forward FFT runs, volk_32fc_s32f_power_spectrum_32f(points, fft, N, N) turns it
 * into dB and the block stops taking input (d_push > 0: whole work() calls are skipped) until get_fft_data has been called, which
 * hands the points out fft-shifted.  One qo_spectrum_work call = one work() call.
 * Not reproducible offline, therefore DEFINED here (parity unpinned for the arithmetic, the buffering / drop logic is pinned to
 * the compiled rx_fft.cpp in tests/test_oracle_ref.py): the DFT is evaluated in double (radix-2, exact twiddles from cos/sin) and
 * rounded to float once; the power spectrum follows VOLK 2.x's generic kernel: re = x.re * (1/N), im likewise (float),
 * 3.01029995663981209120f * log2f(re*re + im*im), an infinite log replaced by -/+127. */
void qo_window_build(int win, int ntaps, float* w) { win_build(win, ntaps, w); }
static void dft_core(double* wr, double* wi, const float* in_c, int N)
{
    int bits = 0; while ((1 << bits) < N) bits++;
    for (int i = 0; i < N; i++) {
        unsigned r = 0; for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1u << (bits - 1 - b);
        wr[r] = (double)in_c[2 * i]; wi[r] = (double)in_c[2 * i + 1];
    }
    for (int len = 2; len <= N; len <<= 1) {
        const int half = len / 2;
        for (int j = 0; j < half; j++) {
            const double a = -2.0 * M_PI * (double)j / (double)len, cr = cos(a), ci = sin(a);
            for (int base = 0; base < N; base += len) {
                const int p = base + j, q = p + half;
                const double tr = wr[q] * cr - wi[q] * ci, ti = wr[q] * ci + wi[q] * cr;
                wr[q] = wr[p] - tr; wi[q] = wi[p] - ti;
                wr[p] += tr; wi[p] += ti;
            }
        }
    }
}
/* the forward DFT as defined for this path: double arithmetic, one rounding to float (also the stand-in for FFTW in oracle/gr_stub) */
void qo_dft_forward(const float* in_c, float* out_c, int n)
{
    double* wr = (double*)malloc(sizeof(double) * (size_t)n); double* wi = (double*)malloc(sizeof(double) * (size_t)n);
    dft_core(wr, wi, in_c, n);
    for (int k = 0; k < n; k++) { out_c[2 * k] = (float)wr[k]; out_c[2 * k + 1] = (float)wi[k]; }
    free(wr); free(wi);
}
struct qo_spectrum {
    int N, win, enabled, data_ready, push;
    unsigned counter;
    float* window; float* buf; float* points;
    double* wr; double* wi;
};
static void spectrum_alloc(qo_spectrum* s)
{
    s->window = (float*)malloc(sizeof(float) * (size_t)s->N); win_build(s->win, s->N, s->window);
    s->buf = (float*)calloc((size_t)s->N * 2, sizeof(float));
    s->points = (float*)calloc((size_t)s->N, sizeof(float));
    s->wr = (double*)malloc(sizeof(double) * (size_t)s->N); s->wi = (double*)malloc(sizeof(double) * (size_t)s->N);
    s->counter = 0; s->data_ready = 0;
}
static void spectrum_free(qo_spectrum* s) { free(s->window); free(s->buf); free(s->points); free(s->wr); free(s->wi); }
qo_spectrum* qo_spectrum_create(int fft_size, int window_type)
{
    if (fft_size < 2 || (fft_size & (fft_size - 1))) return NULL;
    qo_spectrum* s = (qo_spectrum*)calloc(1, sizeof *s);
    s->N = fft_size;
    s->win = (window_type < QO_WIN_HAMMING || window_type > 7) ? QO_WIN_HAMMING : window_type;    /* rx_fft.cpp:176-179 */
    spectrum_alloc(s);
    return s;
}
void qo_spectrum_destroy(qo_spectrum* s) { if (s) { spectrum_free(s); free(s); } }
void qo_spectrum_set_enabled(qo_spectrum* s, int on) { s->enabled = on; }
void qo_spectrum_set_fft_size(qo_spectrum* s, int n)          /* rx_fft.cpp:130-158 */
{
    if (n == s->N || n < 2 || (n & (n - 1))) return;
    spectrum_free(s); s->N = n; spectrum_alloc(s);
}
static void spectrum_execute(qo_spectrum* s)
{
    const int N = s->N;
    dft_core(s->wr, s->wi, s->buf, N);
    const float inorm = 1.0f / (float)N;
    for (int k = 0; k < N; k++) {
        const float re = (float)s->wr[k] * inorm, im = (float)s->wi[k] * inorm;
        float l = log2f(re * re + im * im);
        if (isinf(l)) l = copysignf(127.0f, l);
        s->points[k] = 3.01029995663981209120f * l;
    }
}
void qo_spectrum_work(qo_spectrum* s, const float* iq, long n)
{
    if (s->push > 0 || !s->enabled) return;                     /* rx_fft.cpp:80-84 */
    for (long i = 0; i < n; i++) {
        if (s->counter >= (unsigned)s->N) {
            s->counter = 0;
            spectrum_execute(s);
            s->data_ready = 1; s->push++;
        }
        const float w = s->window[s->counter];
        s->buf[2 * s->counter] = iq[2 * i] * w; s->buf[2 * s->counter + 1] = iq[2 * i + 1] * w;
        s->counter++;
    }
}
int qo_spectrum_get(qo_spectrum* s, float* out)                 /* rx_fft.cpp:112-128 */
{
    s->push = 0;
    if (!s->data_ready) return 0;
    memcpy(out + s->N / 2, s->points, sizeof(float) * (size_t)(s->N / 2));
    memcpy(out, s->points + s->N / 2, sizeof(float) * (size_t)(s->N / 2));
    s->data_ready = 0;
    return s->N;
}

/* ---------------------------------------------------------------------------------------------------------------------------------
 * gr_demod_mmdvm_multi2 / gr_mod_mmdvm_multi2 behind / in front of the polyphase channelizer / synthesizer, ONE channel
 * (/root/reference/src/gr/gr_demod_mmdvm_multi2.cpp:56-126, gr_mod_mmdvm_multi2.cpp:47-126; MMDVM_SAMPLE_RATE = 250000,
 * src/config_mmdvm.h:4: 25 ksps per channelizer port).
 * RX: rational_resampler_ccf(24, 25, low_pass_2(1, 600k, fw, 2000, 60, BH)) -> fft_filter_ccf(low_pass_2(1, 24k, fw, 2000, 60, BH)) ->
 *     rssi_tag_block (an "RSSI" tag every 300 items, rssi_tag_block.cpp:42-63) -> quadrature_demod_cf(24000 / (2 pi 12500)) ->
 *     x 1.0 -> float_to_short(1, 32767) [to gr_mmdvm_sink].
 * TX: short_to_float(1, 32767) -> x 1.0 -> frequency_modulator_fc(2 pi 12500 / 24000) -> fft_filter_ccf(low_pass_2(1, 24k, ...)) ->
 *     x 0.8 -> rational_resampler_ccf(25, 24, low_pass_2(25, 600k, ...)) -> gr_zero_idle_bursts(0) [tags only, not restated here]
 *     [to the synthesizer; behind it x 1 / num_channels -> x bb_gain].
 * float_to_short = volk_32f_s32f_convert_16i (x scale, clip to [-32768, 32767], rintf); short_to_float = volk_16i_s32f_convert_32f in
 * its SIMD form, (float)v * (float)(1.0 / 32767) (the generic kernel divides; an x86 host runs the SIMD one). */
struct qo_mmdvm_rx { int single; resamp_t rs; resamp_t filt; qdemod_t qd; float sum; int nitems; long long n24; float cal; qvec s_a, s_b, s_f, out, rssi_db, rssi_at; };
/* variant 0: one channel of gr_demod_mmdvm_multi2 behind the channelizer (25 ksps in).  variant 1: gr_demod_mmdvm
 * (/root/reference/src/gr/gr_demod_mmdvm.cpp:30-64), the single-channel block at MMDVM_SAMPLE_RATE = 250 ksps: rational_resampler_ccf(12, 125,
 * low_pass_2(12, 12 * 250k, fw, 2000, 60, BH)) -> rssi_tag_block -> low_pass_2(1, 24k, ...) -> quadrature_demod_cf(24000 / (2 pi 10000)) ->
 * x1.0 -> float_to_short: the RSSI tags sit in FRONT of the channel filter there, and the discriminator width is 10 kHz. */
qo_mmdvm_rx* qo_mmdvm_rx_create2(int filter_width, int variant)
{
    static float taps[16384];
    tabs_init();
    qo_mmdvm_rx* r = (qo_mmdvm_rx*)calloc(1, sizeof *r);
    r->single = variant;
    int n;
    if (variant) {
        n = qo_firdes_low_pass_2(12, 12 * 250000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&r->rs, 2, 12, 125, taps, n);
    } else {
        n = qo_firdes_low_pass_2(1, 600000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&r->rs, 2, 24, 25, taps, n);
    }
    n = qo_firdes_low_pass_2(1, 24000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
    resamp_init(&r->filt, 2, 1, 1, taps, n);
    qdemod_init(&r->qd, (float)(24000.0f / (2 * M_PI * (variant ? 10000.0f : 12500.0f))));
    qv_init(&r->s_a, 8); qv_init(&r->s_b, 8); qv_init(&r->s_f, 4); qv_init(&r->out, 2); qv_init(&r->rssi_db, 4); qv_init(&r->rssi_at, 8);
    return r;
}
qo_mmdvm_rx* qo_mmdvm_rx_create(int filter_width) { return qo_mmdvm_rx_create2(filter_width, 0); }
void qo_mmdvm_rx_destroy(qo_mmdvm_rx* r)
{
    if (!r) return;
    resamp_free(&r->rs); resamp_free(&r->filt);
    qv_free(&r->s_a); qv_free(&r->s_b); qv_free(&r->s_f); qv_free(&r->out); qv_free(&r->rssi_db); qv_free(&r->rssi_at);
    free(r);
}
void qo_mmdvm_rx_calibrate_rssi(qo_mmdvm_rx* r, float level) { r->cal = level; }
/* the RSSI tag rule alone (pinned to the compiled rssi_tag_block.cpp in tests/test_oracle_ref.py) */
static void rssi_tag_step(float re, float im, float* sum, int* nitems, float cal, long long at, qvec* db, qvec* where)
{
    const float pwr = re * re + im * im;
    *sum += pwr * pwr;
    *nitems += 1;
    if (*nitems >= 300) {
        const float level = sqrtf(*sum / (float)(*nitems));
        const float v = (float)10.0f * log10f((float)(level + 1.0e-20)) + cal;
        qv_pushf(db, v);
        qv_push(where, &at, 1);
        *sum = 0; *nitems = 0;
    }
}
long qo_rssi_tags_run(const float* in_c, long n, float cal, float* db, long long* at, long cap)
{
    qvec d, w; qv_init(&d, 4); qv_init(&w, 8);
    float sum = 0; int ni = 0;
    for (long i = 0; i < n; i++) rssi_tag_step(in_c[2 * i], in_c[2 * i + 1], &sum, &ni, cal, i, &d, &w);
    const long m = (long)d.n < cap ? (long)d.n : cap;
    memcpy(db, d.d, 4 * (size_t)m); memcpy(at, w.d, 8 * (size_t)m);
    qv_free(&d); qv_free(&w);
    return m;
}
int qo_mmdvm_rx_work(qo_mmdvm_rx* r, const float* iq25k, long n)
{
    r->s_a.n = 0; resamp_work(&r->rs, iq25k, (size_t)n, &r->s_a);
    r->s_b.n = 0; resamp_work(&r->filt, (const float*)r->s_a.d, r->s_a.n, &r->s_b);
    const float* f = (const float*)r->s_b.d;
    const float* g = r->single ? (const float*)r->s_a.d : f;              /* gr_demod_mmdvm tags the resampler's output, multi2 the filter's */
    for (size_t i = 0; i < r->s_b.n; i++) { rssi_tag_step(g[2 * i], g[2 * i + 1], &r->sum, &r->nitems, r->cal, r->n24, &r->rssi_db, &r->rssi_at); r->n24++; }
    r->s_f.n = 0; qdemod_work(&r->qd, f, r->s_b.n, &r->s_f);
    const float* d = (const float*)r->s_f.d;
    for (size_t i = 0; i < r->s_f.n; i++) {
        float v = d[i] * 1.0f;                       /* multiply_const_ff(1.0) */
        v = v * 32767.0f;
        if (v > 32767.0f) v = 32767.0f; else if (v < -32768.0f) v = -32768.0f;
        const short s = (short)rintf(v);
        qv_push(&r->out, &s, 1);
    }
    return 0;
}
long qo_mmdvm_rx_out_items(qo_mmdvm_rx* r) { return (long)r->out.n; }
const short* qo_mmdvm_rx_out_data(qo_mmdvm_rx* r) { return (const short*)r->out.d; }
long qo_mmdvm_rx_rssi_items(qo_mmdvm_rx* r) { return (long)r->rssi_db.n; }
const float* qo_mmdvm_rx_rssi_db(qo_mmdvm_rx* r) { return (const float*)r->rssi_db.d; }
const long long* qo_mmdvm_rx_rssi_at(qo_mmdvm_rx* r) { return (const long long*)r->rssi_at.d; }
void qo_mmdvm_rx_clear(qo_mmdvm_rx* r) { r->out.n = 0; r->rssi_db.n = 0; r->rssi_at.n = 0; }

struct qo_mmdvm_tx { int single; float bb_gain; qo_tx fm; resamp_t filt; resamp_t rs; qvec s_f, s_m, s_b, out; };
/* variant 1: gr_mod_mmdvm (/root/reference/src/gr/gr_mod_mmdvm.cpp:28-70): the same chain up to x0.8, then x bb_gain, then
 * rational_resampler_ccf(125, 12, low_pass_2(125, 125 * 24k, fw, 2000, 60, BH)) to 250 ksps; gr_zero_idle_bursts(0) sits in front of the
 * filter there (behind the resampler in the multi-channel block): qo_mmdvm_tx_zero_samples */
qo_mmdvm_tx* qo_mmdvm_tx_create2(int filter_width, int variant)
{
    static float taps[16384];
    tabs_init();
    qo_mmdvm_tx* t = (qo_mmdvm_tx*)calloc(1, sizeof *t);
    t->single = variant; t->bb_gain = 1.0f;
    t->fm.fm_sens = (float)(2 * M_PI * 12500.0f / 24000.0f);
    int n = qo_firdes_low_pass_2(1, 24000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
    resamp_init(&t->filt, 2, 1, 1, taps, n);
    if (variant) {
        n = qo_firdes_low_pass_2(125, 125 * 24000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->rs, 2, 125, 12, taps, n);
    } else {
        n = qo_firdes_low_pass_2(25, 600000.0, filter_width, 2000, 60, QO_WIN_BLACKMAN_HARRIS, taps, 16384);
        resamp_init(&t->rs, 2, 25, 24, taps, n);
    }
    qv_init(&t->s_f, 4); qv_init(&t->s_m, 8); qv_init(&t->s_b, 8); qv_init(&t->out, 8);
    return t;
}
qo_mmdvm_tx* qo_mmdvm_tx_create(int filter_width) { return qo_mmdvm_tx_create2(filter_width, 0); }
void qo_mmdvm_tx_set_bb_gain(qo_mmdvm_tx* t, float g) { t->bb_gain = g; }
void qo_mmdvm_tx_destroy(qo_mmdvm_tx* t)
{
    if (!t) return;
    resamp_free(&t->filt); resamp_free(&t->rs); qv_free(&t->s_f); qv_free(&t->s_m); qv_free(&t->s_b); qv_free(&t->out);
    free(t->fm.zi_tag_off); free(t->fm.zi_tag_val);
    free(t);
}
int qo_mmdvm_tx_work(qo_mmdvm_tx* t, const short* in, long n)
{
    const float inv = (float)(1.0 / 32767.0);
    t->s_f.n = 0;
    for (long i = 0; i < n; i++) { float v = (float)in[i] * inv; v = v * 1.0f; qv_pushf(&t->s_f, v); }
    t->s_m.n = 0; fm_mod(&t->fm, (const float*)t->s_f.d, t->s_f.n, &t->s_m, 1.0f);
    if (t->single) zero_idle_work(&t->fm, (float*)t->s_m.d, t->s_m.n);        /* gr_mod_mmdvm.cpp:51-58: gr_zero_idle_bursts(0) */
    t->s_b.n = 0; resamp_work(&t->filt, (const float*)t->s_m.d, t->s_m.n, &t->s_b);
    float* m = (float*)t->s_b.d;
    for (size_t i = 0; i < 2 * t->s_b.n; i++) { m[i] = m[i] * 0.8f; if (t->single) m[i] = m[i] * t->bb_gain; }   /* multiply_const_cc(0.8) [, bb_gain] */
    const size_t n_before = t->out.n;
    resamp_work(&t->rs, m, t->s_b.n, &t->out);
    if (!t->single) zero_idle_work(&t->fm, (float*)t->out.d + 2 * n_before, t->out.n - n_before);   /* gr_mod_mmdvm_multi2.cpp:88,108 */
    return 0;
}
/* The "zero_samples" tag (gr_mmdvm_source.cpp:264) as gr_zero_idle_bursts(0) sees it: on item `item_offset` of the block's OWN stream
 * -- the 24 ksps stream behind the FM modulator for gr_mod_mmdvm (every block in front of it is 1:1), the 25 ksps stream behind the
 * x25/24 resampler for gr_mod_mmdvm_multi2 (GNU Radio's scheduler moves a tag across a rate-changing block to
 * floor(offset * 25 / 24 + 1/2); that is runtime behaviour outside /root/reference, so the caller applies it: oracle.mmdvm_tag_item).
 * Same bookkeeping as qo_tx_zero_samples with delay 0. */
int qo_mmdvm_tx_zero_samples(qo_mmdvm_tx* t, long long item_offset, long n_samples)
{
    if (!t || item_offset < 0 || n_samples < 0) return -1;
    qo_tx* z = &t->fm;
    long long start = item_offset;
    uint64_t val = (uint64_t)n_samples;
    if (start < (long long)z->zi_n) {
        const uint64_t late = (uint64_t)((long long)z->zi_n - start);
        if (late >= val) return 0;
        val -= late; start = (long long)z->zi_n;
    }
    for (long i = 0; i < z->zi_ntags; i++) if (z->zi_tag_off[i] == start) return 0;      /* first registered wins */
    if (z->zi_ntags == z->zi_cap) {
        z->zi_cap = z->zi_cap ? 2 * z->zi_cap : 16;
        z->zi_tag_off = (long long*)realloc(z->zi_tag_off, sizeof(long long) * (size_t)z->zi_cap);
        z->zi_tag_val = (uint64_t*)realloc(z->zi_tag_val, sizeof(uint64_t) * (size_t)z->zi_cap);
    }
    z->zi_tag_off[z->zi_ntags] = start; z->zi_tag_val[z->zi_ntags] = val; z->zi_ntags++;
    return 0;
}
long qo_mmdvm_tx_out_items(qo_mmdvm_tx* t) { return (long)t->out.n; }
const float* qo_mmdvm_tx_out_data(qo_mmdvm_tx* t) { return (const float*)t->out.d; }
void qo_mmdvm_tx_clear(qo_mmdvm_tx* t) { t->out.n = 0; }
