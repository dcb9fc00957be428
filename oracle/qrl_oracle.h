/*
 * qrl_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the qradiolink IQ-stream DSP hot path.
 *
 * This is a from-scratch plain-C restatement of the GNU Radio 3.10 blocks that the reference
 * wires together in src/gr/gr_demod_*.cpp / src/gr/gr_mod_*.cpp (see each function's citation in
 * qrl_oracle.c).  It is the parity definition for the CUDA path in qradiolink_b200/.
 *
 * PARITY UNPINNED against real GNU Radio: the reference ships no tests / golden vectors and its
 * arithmetic lives in GNU Radio + VOLK, which are not vendored.  The only piece
 * of the path that compiles from the reference tree itself is src/gr/emphasis.cpp; oracle/_ref pins
 * that one (tests/test_oracle_ref.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product (qradiolink_b200/) never links, imports or calls it.
 */
#ifndef QRL_ORACLE_H
#define QRL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* window ids follow gr::fft::window::win_type */
enum { QO_WIN_HAMMING = 0, QO_WIN_HANN = 1, QO_WIN_BLACKMAN = 2, QO_WIN_RECT = 3, QO_WIN_KAISER = 4,
       QO_WIN_BLACKMAN_HARRIS = 5 };

/* hier-block kinds (one per reference file) */
enum { QO_DEMOD_NBFM = 1, QO_DEMOD_4FSK = 2, QO_DEMOD_QPSK = 3, QO_DEMOD_BPSK = 4, QO_DEMOD_2FSK = 5,
       QO_DEMOD_SSB = 6, QO_DEMOD_AM = 7, QO_DEMOD_GMSK = 8, QO_DEMOD_WBFM = 9, QO_DEMOD_M17 = 10, QO_DEMOD_DMR = 11, QO_DEMOD_DSSS = 12,
       QO_MOD_4FSK = 101, QO_MOD_QPSK = 102, QO_MOD_NBFM = 103, QO_MOD_BPSK = 104, QO_MOD_2FSK = 105,
       QO_MOD_SSB = 106, QO_MOD_GMSK = 107, QO_MOD_M17 = 108, QO_MOD_DMR = 109, QO_MOD_DSSS = 110, QO_MOD_AM = 111 };

/* global numerics switches (used by tests to quantify the documented deviations) */
void qo_set_fir_order(int order);      /* 0 = polyphase/32-lane tree (default, parity order), 1 = sequential oldest-first */
void qo_set_fm_literal(int on);        /* 1 = float phase accumulator + fmodf (GNU Radio literal), 0 = Q32 fixed-point (default) */

/* ---- design functions (firdes etc.) ---- */
int  qo_firdes_low_pass(double gain, double fs, double fc, double tw, int win, float* out, int cap);
int  qo_firdes_low_pass_2(double gain, double fs, double fc, double tw, double att_db, int win, float* out, int cap);
int  qo_firdes_band_pass(double gain, double fs, double lo, double hi, double tw, int win, float* out, int cap);
int  qo_firdes_band_pass_2(double gain, double fs, double lo, double hi, double tw, double att_db, int win, float* out, int cap);
int  qo_firdes_complex_band_pass(double gain, double fs, double lo, double hi, double tw, int win, float* out_c, int cap);
int  qo_firdes_complex_band_pass_2(double gain, double fs, double lo, double hi, double tw, double att_db, int win, float* out_c, int cap);
int  qo_firdes_rrc(double gain, double fs, double symrate, double alpha, int ntaps, float* out, int cap);
int  qo_firdes_gaussian(double gain, double spb, double bt, int ntaps, float* out, int cap);
void qo_deemph_taps(int fs, double tau, double* a2, double* b2);
void qo_preemph_taps(int fs, double tau, double fh, double* a2, double* b2);
void qo_atan_table(float* t257);
void qo_mmse_table(float* t129x8);
void qo_tanh_table(float* t256);
void qo_fxpt_sine_table(float* t1024x2);
void qo_sincosf(float x, float* s, float* c);
float qo_fast_atan2f(float y, float x);
void qo_clock_loop_gains(float loop_bw, float damping, float ted_gain, float* alpha, float* beta);
void qo_control_loop_gains(float loop_bw, float* alpha, float* beta);

/* ---- stand-alone block helpers for unit tests (stateless, zero history) ---- */
/* y[k] = sum_j h[j] x[D k - j], complex x (interleaved), real taps; returns number of outputs */
long qo_fir_decim_ccf(const float* h, int ntaps, int D, const float* x, long n, float* y);
long qo_fir_fff(const float* h, int ntaps, int L, int M, const float* x, long n, float* y, long ycap);
long qo_cc_encode(const uint8_t* bits, long n, uint8_t* out);                 /* streaming K=7 r=1/2 {109,79}, state 0 */
long qo_cc_decode(const uint8_t* soft, long n, uint8_t* out);                 /* fec::decoder(cc_decoder(80,7,2,{109,79})) stream semantics */
void qo_scramble(const uint8_t* in, long n, uint8_t* out);
void qo_descramble(const uint8_t* in, long n, uint8_t* out);

/* ---- RX chains ---- */
typedef struct qo_rx qo_rx;
/* arguments exactly as the reference factories make_gr_demod_*(sps, samp_rate, carrier_freq, filter_width[, fm|sb]) */
qo_rx* qo_rx_create(int kind, int sps, int samp_rate, int carrier_freq, int filter_width, int flag);
void   qo_rx_destroy(qo_rx*);
/* feed T complex samples (interleaved float re,im) of ONE channel */
int    qo_rx_work(qo_rx*, const float* iq, long T);
/* run-time setters of the analog blocks (set_squelch, set_filter_width, set_ctcss(0), set_agc_attack / decay, set_gain); keys as in
 * include/qrl_b200.h; 0 = applied, -1 = this block has no such setter */
enum { QO_PARAM_SQUELCH_DB = 2, QO_PARAM_FILTER_WIDTH = 3, QO_PARAM_CTCSS = 7, QO_PARAM_AGC_ATTACK = 8, QO_PARAM_AGC_DECAY = 9, QO_PARAM_GAIN = 10 };
int    qo_rx_set_param(qo_rx*, int key, double value);
/* gr_demod_base::set_carrier_offset: front-end rotator, phase increment 2*pi*(-offset)/samp_rate */
void   qo_rx_set_carrier_offset(qo_rx*, double offset_hz, double samp_rate);
/* ports: 0 = filtered IQ (complex), 1 = constellation (complex) or audio (float), 2 = bits, 3 = delayed bits */
long   qo_rx_port_items(qo_rx*, int port);
const void* qo_rx_port_data(qo_rx*, int port);
void   qo_rx_port_clear(qo_rx*, int port);
/* debug taps into intermediate streams (tests only): name in {"resamp","demod","rrc","sym","soft"} */
long   qo_rx_dbg_items(qo_rx*, const char* name);
const void* qo_rx_dbg_data(qo_rx*, const char* name);
int    qo_rx_ntaps(qo_rx*, int which, float* out, int cap);

/* ---- front end at device rates >= 2 Msps (gr_demod_base.cpp:1303-1362): rotator at the device rate + /N decimator to 1 Msps ---- */
typedef struct qo_frontend qo_frontend;
qo_frontend* qo_frontend_create(int samp_rate /* multiple of 1e6, >= 2e6 */);
void   qo_frontend_destroy(qo_frontend*);
int    qo_frontend_ntaps(const qo_frontend*);
void   qo_frontend_set_carrier_offset(qo_frontend*, double offset_hz);
long   qo_frontend_work(qo_frontend*, const float* iq, long n, float* out, long cap);

/* ---- TX chains ---- */
typedef struct qo_tx qo_tx;
qo_tx* qo_tx_create(int kind, int sps, int samp_rate, int carrier_freq, int filter_width, int flag);
void   qo_tx_destroy(qo_tx*);
void   qo_tx_set_bb_gain(qo_tx*, float g);
int    qo_tx_set_param(qo_tx*, int key, double value);      /* QO_PARAM_FILTER_WIDTH on QO_MOD_NBFM (gr_mod_nbfm::set_filter_width) */
void   qo_zero_idle_run(const float* in_c, long n, unsigned delay, const long long* tag_item, const long long* tag_val, long ntags, float* out_c);
int    qo_tx_zero_samples(qo_tx*, long long byte_offset, long n_samples);   /* QO_MOD_DMR: the "zero_samples" stream tag */
/* digital: n bytes in; analog (NBFM/SSB): n float audio samples passed as bytes pointer to float */
int    qo_tx_work(qo_tx*, const void* in, long n);
long   qo_tx_out_items(qo_tx*);
const float* qo_tx_out_data(qo_tx*);
void   qo_tx_out_clear(qo_tx*);

/* ---- polyphase channelizer / synthesizer (pfb_channelizer_ccf(M, taps, 1.0) behind stream_to_streams(M);
 *      pfb_synthesizer_ccf(M, taps, false)); streaming state inside the handle ---- */
typedef struct qo_pfb qo_pfb;
qo_pfb* qo_pfb_channelizer_create(int M, const float* taps, int ntaps);
qo_pfb* qo_pfb_synthesizer_create(int M, const float* taps, int ntaps);
void    qo_pfb_destroy(qo_pfb*);
/* x: n complex samples; out: [M][cap] complex, new columns written from column `have`; returns new columns */
long    qo_pfb_channelizer_work(qo_pfb*, const float* x, long n, float* out, long cap, long have);
/* in: [M][stride] complex, n columns; out: n*M complex samples; returns n*M */
long    qo_pfb_synthesizer_work(qo_pfb*, const float* in, long n, long stride, float* out);

/* ---- host-side framing logic (gr_modem.cpp / gr_deframer_bb.cpp restatement) ---- */
/* returns number of frames found; frames written back-to-back (frame_len bytes each) */
long qo_find_frames(const uint8_t* bits, long nbits, uint32_t sync, int sync_bits, int frame_len_bytes,
                    uint8_t* frames, long max_frames);

/* ---- RSSI tap: rssi_block.cpp:25-45 (|x|^2 -> moving_average(2000) -> single_pole_iir(0.04) -> 10 log10 + level) on a
 *      stream of complex samples; returns the value probe_signal_f would hold after the call ---- */
typedef struct qo_rssi qo_rssi;
qo_rssi* qo_rssi_create(float level);
void  qo_rssi_destroy(qo_rssi*);
float qo_rssi_work(qo_rssi*, const float* iq, long n);

/* ---- layer-1 deframer: gr_modem::synchronize / findSync / packBytes restated (gr_modem.cpp:1119-1282).
 *      sync_class 1 = "1K" modes (0xB5), 2 = narrow modes (0xED89 + 24-bit words), 3 = wide modes (IP / video / end).
 *      records: { u32 type, u32 nbytes, payload } of rec_bytes each; returns frames completed by this call ---- */
typedef struct qo_deframer qo_deframer;
qo_deframer* qo_deframer_create(int sync_class, int bit_buf_len, int rx_frame_length);
void  qo_deframer_destroy(qo_deframer*);
long  qo_deframer_work(qo_deframer*, const uint8_t* bits, long n, uint8_t* records, int rec_bytes, long max_frames);
int   qo_deframer_modem_sync(const qo_deframer*);
/* gr_modem::frame (gr_modem.cpp:904-961): sync word (+ burst preamble for IP frames) in front of a payload */
long  qo_frame(const uint8_t* payload, long n, uint32_t frame_type, int one_k_mode, int burst_ip, uint8_t* out, long cap);

/* ---- in-tree reference blocks as stand-alone functions (the chains above call the same code); pinned bit for bit against the
 *      reference sources compiled into oracle/_ref (tests/test_oracle_ref.py) ---- */
void qo_cessb_clipper(const float* in_c, long n, float clip, float* out_c);      /* cessb/clipper_cc_impl.cc:65-95 */
long qo_cessb_stretcher(const float* in_c, long n, float* out_c);                /* cessb/stretcher_cc_impl.cc:70-110, writes n - 2 */
void qo_disc4(const float* m0, const float* m1, const float* m2, const float* m3, long n, float* out_c);   /* gr_4fsk_discriminator.cpp:17-44 */
/* gr_deframer_bb.cpp:83-185 (modem_type 1 / 2 / 3): bit stream in, {sync word bits, following bit_buf_len bits} stream out */
typedef struct qo_dfbb qo_dfbb;
qo_dfbb* qo_dfbb_create(int modem_type);
void  qo_dfbb_destroy(qo_dfbb*);
long  qo_dfbb_work(qo_dfbb*, const uint8_t* bits, long n, uint8_t* out, long cap);

/* rx_fft_c (rx_fft.cpp:44-129): the display spectrum of one stream; one qo_spectrum_work call = one work() call */
void qo_dsss_decoder_taps(const int* code, int code_len, int samples, float* taps_c);
long qo_dsss_decoder_run(const int* code, int code_len, int samples, const float* in_c, long n, long chunk, float* out_c, long cap);
void qo_window_build(int win, int ntaps, float* w);
void qo_dft_forward(const float* in_c, float* out_c, int n);
typedef struct qo_spectrum qo_spectrum;
qo_spectrum* qo_spectrum_create(int fft_size, int window_type);
void qo_spectrum_destroy(qo_spectrum*);
void qo_spectrum_set_enabled(qo_spectrum*, int on);
void qo_spectrum_set_fft_size(qo_spectrum*, int n);
void qo_spectrum_work(qo_spectrum*, const float* iq, long n);
int  qo_spectrum_get(qo_spectrum*, float* out);      /* returns fft_size, or 0 when no spectrum is ready */

/* gr_demod_mmdvm_multi2 / gr_mod_mmdvm_multi2, one channel behind / in front of the polyphase filter bank (25 ksps <-> int16 at 24 ksps) */
typedef struct qo_mmdvm_rx qo_mmdvm_rx;
qo_mmdvm_rx* qo_mmdvm_rx_create(int filter_width);
qo_mmdvm_rx* qo_mmdvm_rx_create2(int filter_width, int variant);      /* variant 1: gr_demod_mmdvm (250 ksps in, x12 / 125) */
void qo_mmdvm_rx_destroy(qo_mmdvm_rx*);
void qo_mmdvm_rx_calibrate_rssi(qo_mmdvm_rx*, float level);
int  qo_mmdvm_rx_work(qo_mmdvm_rx*, const float* iq25k, long n);
long qo_mmdvm_rx_out_items(qo_mmdvm_rx*);
const short* qo_mmdvm_rx_out_data(qo_mmdvm_rx*);
long qo_mmdvm_rx_rssi_items(qo_mmdvm_rx*);
const float* qo_mmdvm_rx_rssi_db(qo_mmdvm_rx*);
const long long* qo_mmdvm_rx_rssi_at(qo_mmdvm_rx*);
void qo_mmdvm_rx_clear(qo_mmdvm_rx*);
long qo_rssi_tags_run(const float* in_c, long n, float cal, float* db, long long* at, long cap);
typedef struct qo_mmdvm_tx qo_mmdvm_tx;
qo_mmdvm_tx* qo_mmdvm_tx_create(int filter_width);
qo_mmdvm_tx* qo_mmdvm_tx_create2(int filter_width, int variant);      /* variant 1: gr_mod_mmdvm (x125 / 12 to 250 ksps, bb_gain) */
void qo_mmdvm_tx_set_bb_gain(qo_mmdvm_tx*, float g);
void qo_mmdvm_tx_destroy(qo_mmdvm_tx*);
int  qo_mmdvm_tx_work(qo_mmdvm_tx*, const short* in, long n);
long qo_mmdvm_tx_out_items(qo_mmdvm_tx*);
const float* qo_mmdvm_tx_out_data(qo_mmdvm_tx*);
void qo_mmdvm_tx_clear(qo_mmdvm_tx*);

#ifdef __cplusplus
}
#endif
#endif
