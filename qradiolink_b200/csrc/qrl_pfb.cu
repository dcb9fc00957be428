// qrl_pfb.cu -- polyphase channelizer / synthesizer, part of libqrl_b200.so.
//
// Replaces gr::filter::pfb_channelizer_ccf(M, taps, 1.0) behind blocks::stream_to_streams(M)
// (/root/reference/src/gr/gr_demod_mmdvm_multi2.cpp:98-107) and gr::filter::pfb_synthesizer_ccf(M, taps, false)
// (/root/reference/src/gr/gr_mod_mmdvm_multi2.cpp:90-92).  The channelizer turns ONE wideband stream into the
// channel-major [M][T/M] layout that qrl_rx_work takes as a device-resident input; the synthesizer is its inverse.
//
//   channelizer: u_k[m] = sum_t taps[k + tM] * x[(m - t)M + (M-1-k)];  out_c[m] = sum_k u_k[m] * exp(+j 2 pi k c / M)
//   synthesizer: v_i[n] = sum_c in_c[n] * exp(+j 2 pi i c / M);        y[nM + i] = sum_t taps[i + tM] * v_i[n - t]
//
// Arithmetic order = the oracle's (oracle/qrl_oracle.c, "polyphase channelizer / synthesizer"): branch FIRs
// accumulate oldest sample first with fmaf; the M-point DFT is the direct sum over k ascending,
// re = fmaf(ur, wr, re); re = fmaf(-ui, wi, re); im = fmaf(ur, wi, im); im = fmaf(ui, wr, im), with twiddles
// (float)cos / (float)sin of the double angle.  Results are bit-identical to the oracle.
//
// Bound: 16 bytes of HBM traffic per wideband sample (8 read + 8 written) against 2*tpf + 4*M FFMA per sample
// (110 for the reference's M = 10, 341-tap prototype): FP32-issue bound on B200, HBM roofline reported beside it.
// There is NO CPU fallback.
#include "../../include/qrl_b200.h"
#include "qrl_tma.cuh"
#include "qrl_handle.hpp"

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

extern "C" int qrl_device_count(void);
void qrl_internal_set_err(const std::string& s);      // qrl_b200.cu: thread-local last error read by qrl_last_error(NULL)

using namespace qrl;

namespace {

// --------------------------------------------------------------------------------------------------------------
// Channelizer, register-tiled instance.  One CTA = TM = G*R consecutive output times of all M channels.
//   stage A: thread (k = tid % M, q = tid / M) holds the TPF taps of branch k in registers and slides over
//            R + TPF - 1 samples of stream M-1-k (stride M in the TMA-fetched window) -> R values of u_k
//   stage B: warp task (c, 32 consecutive m): 10-term DFT out of the transposed u buffer, coalesced 256-byte stores
// R*M mod 16 == 10 keeps the stage-A LDS.64 of a half warp on distinct banks for M = 10.
// The logical input stream is hist ++ x (hist = the previous call's tail, even length so that every window is
// 16-byte aligned for cp.async.bulk).
// --------------------------------------------------------------------------------------------------------------
template <int M, int TPF, int R, int G>
__global__ void __launch_bounds__(M * G)
pfb_chan_kernel(const float* __restrict__ bt /*[M][TPF]*/, const float* __restrict__ w /*[M][2]*/,
                const float2* __restrict__ hist, long long hist_len /* even */, const float2* __restrict__ x, long long n_x,
                long long frame0_e /* x-coordinate of stream sample (m = 0, j = 0): -pending */, long long frames,
                float2* __restrict__ out, long long out_stride, int use_tma)
{
    // x-coordinates: e >= 0 is x[e], e < 0 is hist[hist_len + e] (the previous calls' tail); hist_len is even, so an
    // even e is 16-byte aligned in either buffer and a window that starts on an even e can be fetched by cp.async.bulk
    constexpr int TM = G * R, SPAN = (TM + TPF - 1) * M, SPANP = (SPAN + 3) & ~1, TMP = TM + 1;
    extern __shared__ __align__(128) unsigned char smraw[];
    float2* win = reinterpret_cast<float2*>(smraw);                 // [SPANP]
    float2* ub = win + SPANP;                                       // [M][TMP]
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const long long m0 = static_cast<long long>(blockIdx.x) * TM;
    const long long e0 = frame0_e + (m0 - (TPF - 1)) * M;           // x-coordinate of the first sample stage A reads
    const int woff = static_cast<int>(e0 & 1);
    const long long e0p = e0 - woff;                                // even
    auto fetch = [&](long long e) {
        return e < 0 ? (e >= -hist_len ? hist[hist_len + e] : make_float2(0.0f, 0.0f)) : (e < n_x ? x[e] : make_float2(0.0f, 0.0f));
    };
    if (use_tma) {
        if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
        __syncthreads();
        long long hi = e0p + SPANP;
        if (hi > n_x) hi = n_x & ~1LL;                              // the (short) rest is filled by the loop below
        if (hi < e0p) hi = e0p;
        if (tid == 0) {
            const long long n_h = e0p < 0 ? (hi < 0 ? hi - e0p : -e0p) : 0;
            const long long n_n = hi - e0p - n_h;
            mbar_expect_tx(&bar, static_cast<uint32_t>((hi - e0p) * 8));
            if (n_h > 0) bulk_g2s(win, hist + (hist_len + e0p), static_cast<uint32_t>(n_h * 8), &bar);
            if (n_n > 0) bulk_g2s(win + n_h, x + (e0p + n_h), static_cast<uint32_t>(n_n * 8), &bar);
        }
        for (long long i = hi - e0p + tid; i < SPANP; i += M * G) win[i] = fetch(e0p + i);
        mbar_wait(&bar, 0);
    } else {
        for (int i = tid; i < SPANP; i += M * G) win[i] = fetch(e0p + i);
    }
    __syncthreads();
    // ---- stage A
    {
        // thread -> (branch k, output group q).  The LDS.64 below touch float2 index R M q + (M - 1 - k) + const: a half-warp (16 consecutive
        // tids) is conflict-free when those are distinct mod 16.  With q = tid / M the group stride R M = 90 = 10 (mod 16) collides with
        // the k run of the neighbouring group (2-way conflicts on every load, ncu r02_b: 4 wavefronts instead of 2); visiting the groups
        // in the order 7 q' makes the stride 630 = 6 = -10 (mod 16), i.e. the index is 9 - tid (mod 16): distinct for any 16 consecutive tids.
        const int k = tid % M, qlin = tid / M;
        const int q = (M == 10 && R == 9 && G == 32) ? ((7 * qlin) & 31) : qlin;
        float h[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) h[t] = bt[k * TPF + t];
        float2 acc[R];                                              // (re, im) pairs: one FFMA2 per tap and output
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f);
        const float2* s = win + woff + (q * R) * M + (M - 1 - k);  // s[j*M] = stream sample at time (q*R - (TPF-1) + j)
#pragma unroll
        for (int j = 0; j < R + TPF - 1; j++) {
            const float2 v = s[j * M];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int t = r + TPF - 1 - j;                      // tap index for output r: sample time = m - t
                if (t >= 0 && t < TPF) ffma2(acc[r], h[t], v);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) ub[k * TMP + q * R + r] = acc[r];
    }
    __syncthreads();
    // ---- stage B: warp w owns channel c = w, w + nwarps, ...: its M twiddles stay in registers
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = (M * G) >> 5;
        for (int c = warp; c < M; c += nwarps) {
            float wr[M], wi[M];
#pragma unroll
            for (int k = 0; k < M; k++) { const int qd = (k * c) % M; wr[k] = w[2 * qd]; wi[k] = w[2 * qd + 1]; }
            float2* orow = out + c * out_stride + m0;
            for (int m = lane; m < TM; m += 32) {
                if (m0 + m >= frames) break;
                float2 o = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int k = 0; k < M; k++) {
                    const float2 u = ub[k * TMP + m];
                    // re += u.x wr; re += -u.y wi; im += u.x wi; im += u.y wr  ==  o += u.x (wr, wi); o += u.y (-wi, wr)   (two FFMA2;
                    // fma(-a, b, c) and fma(a, -b, c) are the same operation)
                    ffma2_sp(o, u.x, wr[k], wi[k]);
                    ffma2_sp(o, u.y, -wi[k], wr[k]);
                }
                orow[m] = o;
            }
        }
    }
}

// generic channelizer (any M, tpf): one thread per (m, k) for the branch FIR, then per (m, c) for the DFT
__global__ void __launch_bounds__(256)
pfb_chan_generic_kernel(int M, int tpf, const float* __restrict__ bt, const float* __restrict__ w,
                        const float2* __restrict__ hist, long long hist_len, const float2* __restrict__ x, long long n_x,
                        long long frame0_e, long long frames, float2* __restrict__ out, long long out_stride, int tm)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* ub = reinterpret_cast<float2*>(smraw);                  // [tm][M]
    const long long m0 = static_cast<long long>(blockIdx.x) * tm;
    for (int i = threadIdx.x; i < tm * M; i += blockDim.x) {
        const int k = i % M; const long long m = m0 + i / M;
        float re = 0.0f, im = 0.0f;
        if (m < frames) {
            for (int t = tpf - 1; t >= 0; t--) {
                const long long e = frame0_e + (m - t) * M + (M - 1 - k);
                const float2 v = e < 0 ? (e >= -hist_len ? hist[hist_len + e] : make_float2(0.0f, 0.0f)) : x[e];
                const float h = bt[k * tpf + t];
                re = fmaf(h, v.x, re); im = fmaf(h, v.y, im);
            }
        }
        ub[i] = make_float2(re, im);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tm * M; i += blockDim.x) {
        const int ml = i % tm, c = i / tm;
        if (m0 + ml >= frames) continue;
        float re = 0.0f, im = 0.0f;
        for (int k = 0; k < M; k++) {
            const int qd = static_cast<int>((static_cast<long long>(k) * c) % M);
            const float wr = w[2 * qd], wi = w[2 * qd + 1];
            const float2 u = ub[ml * M + k];
            re = fmaf(u.x, wr, re); re = fmaf(-u.y, wi, re);
            im = fmaf(u.x, wi, im); im = fmaf(u.y, wr, im);
        }
        out[c * out_stride + m0 + ml] = make_float2(re, im);
    }
}

// the last new_len samples of the stream (x-coordinates n_x - new_len .. n_x - 1) -> the other history buffer
__global__ void pfb_tail_kernel(const float2* __restrict__ hist, long long hist_len, const float2* __restrict__ x,
                                long long n_x, float2* __restrict__ dst, long long new_len)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= new_len) return;
    const long long e = n_x - new_len + i;
    dst[i] = e < 0 ? (e >= -hist_len ? hist[hist_len + e] : make_float2(0.0f, 0.0f)) : x[e];
}

// --------------------------------------------------------------------------------------------------------------
// Synthesizer.  One CTA = TN = G*R consecutive input times n of all M channels.
//   stage A: thread (n lane-fastest, i): v_i[n] for n0 - (TPF-1) .. n0 + TN - 1 out of the [M][cols] input window
//            (the columns before the call come from the per-channel history rows)
//   stage B: thread (i = tid % M, q = tid / M): R outputs y[(n0 + qR + r) M + i], taps of branch i in registers,
//            sliding over v_i (contiguous in n); consecutive lanes write consecutive wideband samples
// --------------------------------------------------------------------------------------------------------------
template <int M, int TPF, int R, int G>
__global__ void __launch_bounds__(M * G)
pfb_synth_kernel(const float* __restrict__ bt, const float* __restrict__ w,
                 const float2* __restrict__ hist /*[M][TPF-1]*/, const float2* __restrict__ in, long long in_stride,
                 long long n_cols, float2* __restrict__ out)
{
    constexpr int TN = G * R, COLS = TN + TPF - 1, VP = COLS | 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* xin = reinterpret_cast<float2*>(smraw);                 // [M][COLS]
    float2* vb = xin + M * COLS;                                    // [M][VP]
    const int tid = threadIdx.x;
    const long long n0 = static_cast<long long>(blockIdx.x) * TN;
    for (int i = tid; i < M * COLS; i += M * G) {
        const int c = i / COLS, col = i % COLS;
        const long long n = n0 - (TPF - 1) + col;
        xin[i] = n < 0 ? hist[c * (TPF - 1) + (TPF - 1) + n] : (n < n_cols ? in[c * in_stride + n] : make_float2(0.0f, 0.0f));
    }
    __syncthreads();
    {
        const int warp = tid >> 5, lane = tid & 31, nwarps = (M * G) >> 5;
        for (int br = warp; br < M; br += nwarps) {              // warp owns branch br: twiddles in registers
            float wr[M], wi[M];
#pragma unroll
            for (int c = 0; c < M; c++) { const int qd = (br * c) % M; wr[c] = w[2 * qd]; wi[c] = w[2 * qd + 1]; }
            for (int col = lane; col < COLS; col += 32) {
                float2 o = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int c = 0; c < M; c++) {
                    const float2 u = xin[c * COLS + col];
                    ffma2_sp(o, u.x, wr[c], wi[c]);             // re += u.x wr, im += u.x wi
                    ffma2_sp(o, u.y, -wi[c], wr[c]);            // re += -u.y wi, im += u.y wr
                }
                vb[br * VP + col] = o;
            }
        }
    }
    __syncthreads();
    {
        const int br = tid % M, q = tid / M;
        float h[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) h[t] = bt[br * TPF + t];
        float2 acc[R];                                              // (re, im) pairs: one FFMA2 per tap and output
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f);
        const float2* s = vb + br * VP + q * R;                     // s[j] = v_br[n0 + qR - (TPF-1) + j]
#pragma unroll
        for (int j = 0; j < R + TPF - 1; j++) {
            const float2 v = s[j];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int t = r + TPF - 1 - j;
                if (t >= 0 && t < TPF) ffma2(acc[r], h[t], v);
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const long long n = n0 + q * R + r;
            if (n < n_cols) out[n * M + br] = acc[r];
        }
    }
}

__global__ void __launch_bounds__(256)
pfb_synth_generic_kernel(int M, int tpf, const float* __restrict__ bt, const float* __restrict__ w,
                         const float2* __restrict__ hist, const float2* __restrict__ in, long long in_stride,
                         long long n_cols, float2* __restrict__ out, int tn)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    float2* vb = reinterpret_cast<float2*>(smraw);                  // [M][tn + tpf - 1]
    const int cols = tn + tpf - 1;
    const long long n0 = static_cast<long long>(blockIdx.x) * tn;
    for (int i = threadIdx.x; i < M * cols; i += blockDim.x) {
        const int br = i / cols, col = i % cols;
        const long long n = n0 - (tpf - 1) + col;
        float re = 0.0f, im = 0.0f;
        for (int c = 0; c < M; c++) {
            const float2 u = n < 0 ? hist[c * (tpf - 1) + (tpf - 1) + n] : (n < n_cols ? in[c * in_stride + n] : make_float2(0.0f, 0.0f));
            const int qd = static_cast<int>((static_cast<long long>(br) * c) % M);
            const float wr = w[2 * qd], wi = w[2 * qd + 1];
            re = fmaf(u.x, wr, re); re = fmaf(-u.y, wi, re);
            im = fmaf(u.x, wi, im); im = fmaf(u.y, wr, im);
        }
        vb[i] = make_float2(re, im);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tn * M; i += blockDim.x) {
        const int br = i % M, nl = i / M;
        const long long n = n0 + nl;
        if (n >= n_cols) continue;
        float re = 0.0f, im = 0.0f;
        for (int t = tpf - 1; t >= 0; t--) {
            const float2 v = vb[br * cols + nl + (tpf - 1) - t];
            const float h = bt[br * tpf + t];
            re = fmaf(h, v.x, re); im = fmaf(h, v.y, im);
        }
        out[n * M + br] = make_float2(re, im);
    }
}

// new per-channel history rows: the last tpf-1 input columns of (hist ++ in)
__global__ void pfb_synth_tail_kernel(int M, int hl, const float2* __restrict__ hist, const float2* __restrict__ in,
                                      long long in_stride, long long n_cols, float2* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * hl) return;
    const int c = i / hl, j = i % hl;
    const long long n = n_cols - hl + j;
    dst[i] = n < 0 ? hist[c * hl + hl + n] : in[c * in_stride + n];
}

#define CKP(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                            \
            qrl_internal_set_err(h->err);                                                            \
            return QRL_ECUDA;                                                                        \
        }                                                                                            \
    } while (0)

}  // namespace

struct qrl_pfb : QrlHandleBase {        // err / device / stream / launches / allocs: qrl_handle.hpp
    int kind = 0, M = 0, tpf = 0;
    long max_in = 0;
    float *d_bt = nullptr, *d_w = nullptr;
    // channelizer: tail of the stream = [pad][(tpf-1)*M history][pend < M pending samples], hist_len even
    // synthesizer: [M][tpf-1] last input columns
    float2* d_hist[2] = { nullptr, nullptr }; int hist_cur = 0; long long hist_len = 0; int pend = 0;
    float2* d_stage = nullptr;          // staging for host input
    float2* d_out = nullptr; long long out_stride = 0; long long out_items = 0;
    bool fast = false;
};

namespace {
template <class T> int pfb_alloc(qrl_pfb* h, T** p, size_t n, bool zero = true)
{
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(n * sizeof(T), 16);
    if (cudaMalloc(&q, bytes) != cudaSuccess) { h->err = "cudaMalloc failed"; qrl_internal_set_err(h->err); return QRL_ENOMEM; }
    if (zero && cudaMemsetAsync(q, 0, bytes, h->stream) != cudaSuccess) { h->err = "cudaMemset failed"; qrl_internal_set_err(h->err); return QRL_ECUDA; }
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return QRL_OK;
}
constexpr int kFastM = 10, kFastTpf = 35, kChanR = 9, kChanG = 32, kSynR = 10, kSynG = 32;
long long chan_hist_len(int tpf, int M, int pend) { return (static_cast<long long>(tpf - 1) * M + pend + 2) & ~1LL; }
}  // namespace

extern "C" {

int qrl_pfb_destroy(qrl_pfb* h)
{
    if (!h) return QRL_EINVAL;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return QRL_OK;
}

int qrl_pfb_create(int kind, int M, const float* taps, int ntaps, long max_in, int device, qrl_pfb** out)
{
    if (!out || !taps || (kind != QRL_PFB_CHANNELIZER && kind != QRL_PFB_SYNTHESIZER) || M < 1 || M > 64 || ntaps < 1 ||
        max_in < 1 || (ntaps + M - 1) / M > 128) {
        qrl_internal_set_err("qrl_pfb_create: bad argument");
        return QRL_EINVAL;
    }
    *out = nullptr;
    if (qrl_device_count() <= device) { qrl_internal_set_err("qrl_pfb_create: no CUDA device (this library has no CPU fallback)"); return QRL_ENODEV; }
    qrl_pfb* h = new qrl_pfb();
    h->kind = kind; h->M = M; h->tpf = (ntaps + M - 1) / M; h->max_in = max_in; h->device = device;
    auto fail = [&](int rc) { std::string e = h->err; qrl_pfb_destroy(h); qrl_internal_set_err(e); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(QRL_ECUDA); }
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { h->err = "stream create failed"; return fail(QRL_ECUDA); }
    h->own_stream = true;
    // branch taps [M][tpf] (polyphase_filterbank::set_taps: filter k holds taps[k + t*M], zero padded) and twiddles
    std::vector<float> bt(static_cast<size_t>(M) * h->tpf, 0.0f), w(static_cast<size_t>(M) * 2);
    for (int j = 0; j < ntaps; j++) bt[static_cast<size_t>(j % M) * h->tpf + j / M] = taps[j];
    for (int q = 0; q < M; q++) {
        const double a = 2.0 * 3.14159265358979323846 * static_cast<double>(q) / static_cast<double>(M);
        w[2 * q] = static_cast<float>(std::cos(a)); w[2 * q + 1] = static_cast<float>(std::sin(a));
    }
    int rc;
    if ((rc = pfb_alloc(h, &h->d_bt, bt.size(), false)) || (rc = pfb_alloc(h, &h->d_w, w.size(), false))) return fail(rc);
    if (cudaMemcpyAsync(h->d_bt, bt.data(), bt.size() * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(h->d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { h->err = "tap upload failed"; return fail(QRL_ECUDA); }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "tap upload sync failed"; return fail(QRL_ECUDA); }
    h->fast = (M == kFastM && h->tpf == kFastTpf);
    if (kind == QRL_PFB_CHANNELIZER) {
        const size_t hcap = static_cast<size_t>(h->tpf) * M + 4;
        if ((rc = pfb_alloc(h, &h->d_hist[0], hcap)) || (rc = pfb_alloc(h, &h->d_hist[1], hcap))) return fail(rc);
        h->pend = 0; h->hist_len = chan_hist_len(h->tpf, M, 0);
        h->out_stride = (max_in + M - 1) / M + 2;
        if ((rc = pfb_alloc(h, &h->d_out, static_cast<size_t>(h->out_stride) * M))) return fail(rc);
        if ((rc = pfb_alloc(h, &h->d_stage, static_cast<size_t>(max_in)))) return fail(rc);
    } else {
        const size_t hl = static_cast<size_t>(std::max(1, h->tpf - 1)) * M;
        if ((rc = pfb_alloc(h, &h->d_hist[0], hl)) || (rc = pfb_alloc(h, &h->d_hist[1], hl))) return fail(rc);
        h->out_stride = max_in * M;
        if ((rc = pfb_alloc(h, &h->d_out, static_cast<size_t>(max_in) * M))) return fail(rc);
        if ((rc = pfb_alloc(h, &h->d_stage, static_cast<size_t>(max_in) * M))) return fail(rc);
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "create sync failed"; return fail(QRL_ECUDA); }
    *out = h;
    return QRL_OK;
}

int qrl_pfb_set_stream(qrl_pfb* h, void* cuda_stream)
{
    if (!h) return QRL_EINVAL;
    if (!cuda_stream) return QRL_OK;
    CKP(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = static_cast<cudaStream_t>(cuda_stream); h->own_stream = false;
    return QRL_OK;
}

int qrl_pfb_work(qrl_pfb* h, const void* in, long n_in, long in_stride, int in_on_device, long* n_out)
{
    if (!h || (!in && n_in > 0) || n_in < 0 || n_in > h->max_in) { qrl_internal_set_err("qrl_pfb_work: bad argument"); return QRL_EINVAL; }
    CKP(cudaSetDevice(h->device));
    const int M = h->M, tpf = h->tpf;
    const float2* hist = h->d_hist[h->hist_cur];
    float2* hist_next = h->d_hist[h->hist_cur ^ 1];
    if (h->kind == QRL_PFB_CHANNELIZER) {
        const float2* x = static_cast<const float2*>(in);
        if (!in_on_device && n_in > 0) {
            CKP(cudaMemcpyAsync(h->d_stage, in, sizeof(float2) * n_in, cudaMemcpyHostToDevice, h->stream));
            x = h->d_stage;
        }
        const long long frames = (static_cast<long long>(h->pend) + n_in) / M;
        if (frames > 0) {
            if (h->fast) {
                constexpr int TM = kChanR * kChanG, SPAN = (TM + kFastTpf - 1) * kFastM, SPANP = (SPAN + 3) & ~1;
                const size_t smem = sizeof(float2) * (SPANP + kFastM * (TM + 1));
                auto kern = pfb_chan_kernel<kFastM, kFastTpf, kChanR, kChanG>;
                static bool attr[16] = { false };    // per device: function attributes belong to the device's context
                if (!attr[h->device & 15]) { CKP(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); attr[h->device & 15] = true; }
                const int use_tma = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
                kern<<<static_cast<unsigned>((frames + TM - 1) / TM), kFastM * kChanG, smem, h->stream>>>(
                    h->d_bt, h->d_w, hist, h->hist_len, x, n_in, -static_cast<long long>(h->pend), frames,
                    h->d_out, h->out_stride, use_tma);
            } else {
                const int tm = std::max(1, 2048 / M);
                pfb_chan_generic_kernel<<<static_cast<unsigned>((frames + tm - 1) / tm), 256, sizeof(float2) * tm * M, h->stream>>>(
                    M, tpf, h->d_bt, h->d_w, hist, h->hist_len, x, n_in, -static_cast<long long>(h->pend), frames,
                    h->d_out, h->out_stride, tm);
            }
            h->launches++;
        }
        const int new_pend = static_cast<int>((static_cast<long long>(h->pend) + n_in) % M);
        const long long new_len = chan_hist_len(tpf, M, new_pend);
        if (n_in > 0) {
            pfb_tail_kernel<<<static_cast<unsigned>((new_len + 255) / 256), 256, 0, h->stream>>>(hist, h->hist_len, x, n_in, hist_next, new_len);
            h->launches++;
            h->hist_cur ^= 1; h->hist_len = new_len; h->pend = new_pend;
        }
        h->out_items = frames;
        if (n_out) *n_out = static_cast<long>(frames);
    } else {
        const float2* x = static_cast<const float2*>(in);
        long long stride = in_stride;
        if (!in_on_device && n_in > 0) {
            CKP(cudaMemcpy2DAsync(h->d_stage, sizeof(float2) * n_in, in, sizeof(float2) * in_stride, sizeof(float2) * n_in, M,
                                  cudaMemcpyHostToDevice, h->stream));
            x = h->d_stage; stride = n_in;
        }
        if (n_in > 0) {
            const int hl = tpf - 1;
            if (h->fast) {
                constexpr int TN = kSynR * kSynG, COLS = TN + kFastTpf - 1, VP = COLS | 1;
                const size_t smem = sizeof(float2) * (kFastM * COLS + kFastM * VP);
                auto kern = pfb_synth_kernel<kFastM, kFastTpf, kSynR, kSynG>;
                static bool attr[16] = { false };    // per device: function attributes belong to the device's context
                if (!attr[h->device & 15]) { CKP(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); attr[h->device & 15] = true; }
                kern<<<static_cast<unsigned>((n_in + TN - 1) / TN), kFastM * kSynG, smem, h->stream>>>(
                    h->d_bt, h->d_w, hist, x, stride, n_in, h->d_out);
            } else {
                const int tn = std::max(1, 2048 / M);
                pfb_synth_generic_kernel<<<static_cast<unsigned>((n_in + tn - 1) / tn), 256, sizeof(float2) * M * (tn + tpf - 1), h->stream>>>(
                    M, tpf, h->d_bt, h->d_w, hist, x, stride, n_in, h->d_out, tn);
            }
            h->launches++;
            if (hl > 0) {
                pfb_synth_tail_kernel<<<(M * hl + 255) / 256, 256, 0, h->stream>>>(M, hl, hist, x, stride, n_in, hist_next);
                h->launches++;
                h->hist_cur ^= 1;
            }
        }
        h->out_items = static_cast<long long>(n_in) * M;
        if (n_out) *n_out = static_cast<long>(h->out_items);
    }
    CKP(cudaGetLastError());
    return QRL_OK;
}

int qrl_pfb_sync(qrl_pfb* h)
{
    if (!h) return QRL_EINVAL;
    CKP(cudaStreamSynchronize(h->stream));
    return QRL_OK;
}

int qrl_pfb_out_device(qrl_pfb* h, void** data, long* stride, long* items)
{
    if (!h) return QRL_EINVAL;
    if (data) *data = h->d_out;
    if (stride) *stride = static_cast<long>(h->out_stride);
    if (items) *items = static_cast<long>(h->out_items);
    return QRL_OK;
}

int qrl_pfb_read(qrl_pfb* h, void* host_dst, long dst_stride)
{
    if (!h || !host_dst) return QRL_EINVAL;
    CKP(cudaSetDevice(h->device));
    if (h->out_items > 0) {
        if (h->kind == QRL_PFB_CHANNELIZER)
            CKP(cudaMemcpy2DAsync(host_dst, sizeof(float2) * dst_stride, h->d_out, sizeof(float2) * h->out_stride,
                                  sizeof(float2) * h->out_items, h->M, cudaMemcpyDeviceToHost, h->stream));
        else
            CKP(cudaMemcpyAsync(host_dst, h->d_out, sizeof(float2) * h->out_items, cudaMemcpyDeviceToHost, h->stream));
    }
    CKP(cudaStreamSynchronize(h->stream));
    return QRL_OK;
}

long qrl_pfb_launch_count(qrl_pfb* h) { return h ? h->launches : 0; }

}  // extern "C"
