// C-ABI shim over the reference's IN-TREE GNU Radio blocks, compiled UNMODIFIED from /root/reference/src/gr against the
// minimal runtime stand-in in oracle/gr_stub/ (no scheduler: this file plays the scheduler and calls work() /
// general_work() with the buffers a GNU Radio scheduler would present: history in front, output multiples respected).
// TEST INFRASTRUCTURE ONLY (oracle/_ref/libqrl_ref_blocks.so): pins oracle/qrl_oracle.c's restatements of
//   gr_4fsk_discriminator.cpp, cessb/clipper_cc_impl.cc, cessb/stretcher_cc_impl.cc, gr_deframer_bb.cpp,
//   gr_bit_sink.cpp, gr_audio_sink.cpp, gr_const_sink.cpp, dsss_encoder_bb_impl.cc, dsss_decoder_cc_impl.cc
// to the reference's own code (tests/test_oracle_ref.py).  Nothing of the product links or loads this.
#include <gnuradio/block.h>

#include "gr_4fsk_discriminator.h"
#include "gr_deframer_bb.h"
#include "gr_bit_sink.h"
#include "gr_audio_sink.h"
#include "gr_const_sink.h"
#include "cessb/clipper_cc.h"
#include "cessb/stretcher_cc.h"
#include "dsss_encoder_bb_impl.h"
#include "dsss_decoder_cc_impl.h"

namespace {
struct Any { std::shared_ptr<gr::block> b; };
template <class T> T* as(void* h) { return dynamic_cast<T*>(static_cast<Any*>(h)->b.get()); }
}

extern "C" {

void ref_block_destroy(void* h) { delete static_cast<Any*>(h); }

// ---------------------------------------------------------------- gr_4fsk_discriminator
void ref_disc4(const float* m0, const float* m1, const float* m2, const float* m3, long n, float* out_c)
{
    auto blk = make_gr_4fsk_discriminator();
    gr_vector_const_void_star in = { m0, m1, m2, m3 };
    gr_vector_void_star out = { out_c };
    blk->work(static_cast<int>(n), in, out);
}

// ---------------------------------------------------------------- cessb::clipper_cc (sync block, output multiple 1024)
long ref_cessb_clipper(const float* in_c, long n, float clip, float* out_c)
{
    auto blk = gr::cessb::clipper_cc::make(clip);
    const long n_out = n / 1024 * 1024;
    gr_vector_const_void_star in = { in_c };
    gr_vector_void_star out = { out_c };
    if (n_out > 0) blk->work(static_cast<int>(n_out), in, out);
    return n_out;
}

// ---------------------------------------------------------------- cessb::stretcher_cc (general block, forecast = n + 2 per chunk)
// fed chunk by chunk like the scheduler would (consumed items drop off the front), so the carried d_env state is exercised
long ref_cessb_stretcher(const float* in_c, long n, long chunk_items, float* out_c)
{
    auto blk = gr::cessb::stretcher_cc::make();
    long done = 0;
    if (chunk_items < 1024) chunk_items = 1024;
    chunk_items = chunk_items / 1024 * 1024;
    while (true) {
        long want = chunk_items;
        while (want >= 1024) {
            gr_vector_int req(1, 0);
            blk->forecast(static_cast<int>(want), req);
            if (done + req[0] <= n) break;
            want -= 1024;
        }
        if (want < 1024) break;
        gr_vector_int nin(1, static_cast<int>(n - done));
        gr_vector_const_void_star in = { in_c + 2 * done };
        gr_vector_void_star out = { out_c + 2 * done };
        const int produced = blk->general_work(static_cast<int>(want), nin, in, out);
        const long consumed = blk->take_consumed();
        if (produced != want || consumed != want) return -1;
        done += want;
    }
    return done;
}

// ---------------------------------------------------------------- gr_deframer_bb
void* ref_dfbb_create(int modem_type) { return new Any{ make_gr_deframer_bb(modem_type) }; }
long ref_dfbb_work(void* h, const unsigned char* bits, long n, unsigned char* out, long cap)
{
    auto* d = as<gr_deframer_bb>(h);
    gr_vector_const_void_star in = { bits };
    gr_vector_void_star outv;
    d->work(static_cast<int>(n), in, outv);
    std::vector<unsigned char>* v = d->get_data();
    const long m = static_cast<long>(v->size());
    for (long i = 0; i < m && i < cap; i++) out[i] = (*v)[i];
    delete v;
    return m < cap ? m : cap;
}

// ---------------------------------------------------------------- sinks: work() appends, get_data() -> packet or nullptr (-1)
void* ref_bit_sink_create() { return new Any{ make_gr_bit_sink() }; }
int ref_bit_sink_work(void* h, const unsigned char* in, int n)
{
    gr_vector_const_void_star iv = { in }; gr_vector_void_star ov;
    return as<gr_bit_sink>(h)->work(n, iv, ov);
}
long ref_bit_sink_get(void* h, unsigned char* out, long cap)
{
    std::vector<unsigned char>* v = as<gr_bit_sink>(h)->get_data();
    if (!v) return -1;
    const long m = static_cast<long>(v->size());
    for (long i = 0; i < m && i < cap; i++) out[i] = (*v)[i];
    delete v;
    return m;
}
void ref_bit_sink_flush(void* h) { as<gr_bit_sink>(h)->flush(); }

void* ref_audio_sink_create() { return new Any{ make_gr_audio_sink() }; }
int ref_audio_sink_work(void* h, const float* in, int n)
{
    gr_vector_const_void_star iv = { in }; gr_vector_void_star ov;
    return as<gr_audio_sink>(h)->work(n, iv, ov);
}
long ref_audio_sink_get(void* h, float* out, long cap)
{
    std::vector<float>* v = as<gr_audio_sink>(h)->get_data();
    if (!v) return -1;
    const long m = static_cast<long>(v->size());
    for (long i = 0; i < m && i < cap; i++) out[i] = (*v)[i];
    delete v;
    return m;
}

void* ref_const_sink_create() { return new Any{ make_gr_const_sink() }; }
int ref_const_sink_work(void* h, const float* in_c, int n)
{
    gr_vector_const_void_star iv = { in_c }; gr_vector_void_star ov;
    return as<gr_const_sink>(h)->work(n, iv, ov);
}
long ref_const_sink_get(void* h, float* out_c, long cap)
{
    std::vector<gr_complex>* v = as<gr_const_sink>(h)->get_data();
    if (!v) return -1;
    const long m = static_cast<long>(v->size());
    for (long i = 0; i < m && i < cap; i++) { out_c[2 * i] = (*v)[i].real(); out_c[2 * i + 1] = (*v)[i].imag(); }
    delete v;
    return m;
}

// ---------------------------------------------------------------- gr::dsss::dsss_encoder_bb (bytes -> chips, 8 * code length per byte)
long ref_dsss_encode(const int* code, int code_len, const unsigned char* bytes, long n, unsigned char* chips)
{
    auto blk = gr::dsss::dsss_encoder_bb::make(std::vector<int>(code, code + code_len));
    gr_vector_int nin(1, static_cast<int>(n));
    gr_vector_const_void_star in = { bytes };
    gr_vector_void_star out = { chips };
    const int produced = blk->general_work(static_cast<int>(n * 8 * code_len), nin, in, out);
    blk->take_consumed();
    return produced;
}

// ---------------------------------------------------------------- gr::dsss::dsss_decoder_cc
// matched-filter taps the constructor builds (RRC-shaped reversed code), and general_work over a buffer that starts
// `hist` items before the first new item, exactly as the scheduler lays it out (history() - 1 old items in front).
void* ref_dsss_decoder_create(const int* code, int code_len, float sps) { return new Any{ gr::dsss::dsss_decoder_cc::make(std::vector<int>(code, code + code_len), sps) }; }
int ref_dsss_decoder_taps(void* h, float* out_c, int cap)
{
    auto t = as<gr::dsss::dsss_decoder_cc>(h)->taps();
    for (int i = 0; i < static_cast<int>(t.size()) && i < cap; i++) { out_c[2 * i] = t[i].real(); out_c[2 * i + 1] = t[i].imag(); }
    return static_cast<int>(t.size());
}
int ref_dsss_decoder_history(void* h) { return static_cast<int>(as<gr::dsss::dsss_decoder_cc>(h)->history()); }
// in_c points at the first NEW item; the caller guarantees `before` readable items in front of it and `after` items from it on
long ref_dsss_decoder_work(void* h, const float* in_c, int noutput, float* out_c, long* consumed)
{
    auto* d = as<gr::dsss::dsss_decoder_cc>(h);
    gr_vector_int nin(1, 0);
    gr_vector_const_void_star in = { in_c };
    gr_vector_void_star out = { out_c };
    const int produced = d->general_work(noutput, nin, in, out);
    *consumed = d->take_consumed();
    return produced;
}

}  // extern "C"

// ---------------------------------------------------------------- gr_zero_idle_bursts (sync block with history 2*SAMPLES_PER_SLOT, stream tags)
// Fed chunk by chunk like the scheduler: every work() sees history()-1 old items in front of the new ones (zeros at the start of
// the stream) and the "zero_samples" tags whose absolute offset falls into the window.
#include "gr_zero_idle_bursts.h"
extern "C" long ref_zero_idle(const float* in_c, long n, unsigned delay, const long long* tag_off, const long long* tag_val, long ntags,
                              const long* chunks, long nchunks, float* out_c)
{
    auto blk = make_gr_zero_idle_bursts(delay);
    for (long i = 0; i < ntags; i++) {
        gr::tag_t t; t.offset = static_cast<uint64_t>(tag_off[i]); t.key = pmt::string_to_symbol("zero_samples");
        t.value = pmt::from_uint64(static_cast<uint64_t>(tag_val[i]));
        blk->stub_add_input_tag(t);
    }
    const long H = static_cast<long>(blk->history()) - 1;
    std::vector<gr_complex> buf(static_cast<size_t>(H + n), gr_complex(0, 0));
    std::memcpy(buf.data() + H, in_c, sizeof(gr_complex) * static_cast<size_t>(n));
    long done = 0;
    for (long k = 0; k < nchunks && done < n; k++) {
        const long m = std::min(chunks[k], n - done);
        gr_vector_const_void_star in = { buf.data() + done };
        gr_vector_void_star out = { reinterpret_cast<gr_complex*>(out_c) + done };
        blk->work(static_cast<int>(m), in, out);
        blk->stub_advance(static_cast<uint64_t>(m), static_cast<uint64_t>(m));
        done += m;
    }
    return done;
}

// ---------------------------------------------------------------- rx_fft_c (display spectrum; FFTW replaced by the oracle's DFT, see gr_stub/gnuradio/fft/fft.h)
#include "rx_fft.h"
extern "C" {
void* ref_rx_fft_create(unsigned fftsize, int wintype) { auto* a = new Any; a->b = make_rx_fft_c(fftsize, wintype); return a; }
void ref_rx_fft_set_enabled(void* h, int on) { as<rx_fft_c>(h)->set_enabled(on != 0); }
void ref_rx_fft_set_fft_size(void* h, unsigned n) { as<rx_fft_c>(h)->set_fft_size(n); }
void ref_rx_fft_work(void* h, const float* in_c, int n)
{
    gr_vector_const_void_star in = { in_c };
    gr_vector_void_star out;
    as<rx_fft_c>(h)->work(n, in, out);
}
unsigned ref_rx_fft_get(void* h, float* points) { unsigned n = 0; as<rx_fft_c>(h)->get_fft_data(points, n); return n; }
}

// ---------------------------------------------------------------- gr_sample_sink (time-domain display tap: window, enable, 524288-item drop rule)
#include "gr_sample_sink.h"
extern "C" {
void* ref_sample_sink_create() { return new Any{ make_gr_sample_sink() }; }
void ref_sample_sink_set_enabled(void* h, int on) { as<gr_sample_sink>(h)->set_enabled(on != 0); }
void ref_sample_sink_set_window(void* h, unsigned n) { as<gr_sample_sink>(h)->set_sample_window(n); }
int ref_sample_sink_work(void* h, const float* in_c, int n)
{
    gr_vector_const_void_star iv = { in_c }; gr_vector_void_star ov;
    return as<gr_sample_sink>(h)->work(n, iv, ov);
}
long ref_sample_sink_get(void* h, float* out_c, long cap)
{
    std::vector<gr_complex>* v = as<gr_sample_sink>(h)->get_data();
    if (!v) return -1;
    const long m = static_cast<long>(v->size());
    for (long i = 0; i < m && i < cap; i++) { out_c[2 * i] = (*v)[i].real(); out_c[2 * i + 1] = (*v)[i].imag(); }
    delete v;
    return m;
}
}

// ---------------------------------------------------------------- rssi_tag_block (an "RSSI" stream tag every 300 items)
#include "rssi_tag_block.h"
extern "C" long ref_rssi_tags(const float* in_c, long n, float cal, const long* chunks, long nchunks, float* db, long long* at, long cap)
{
    auto blk = make_rssi_tag_block();
    blk->calibrate_rssi(cal);
    std::vector<gr_complex> out(static_cast<size_t>(n));
    long done = 0;
    for (long k = 0; k < nchunks && done < n; k++) {
        const long m = std::min(chunks[k], n - done);
        gr_vector_const_void_star in = { reinterpret_cast<const gr_complex*>(in_c) + done };
        gr_vector_void_star o = { out.data() + done };
        blk->work(static_cast<int>(m), in, o);
        blk->stub_advance(static_cast<uint64_t>(m), static_cast<uint64_t>(m));
        done += m;
    }
    long cnt = 0;
    for (const auto& t : blk->stub_out_tags()) if (cnt < cap) { db[cnt] = pmt::to_float(t.value); at[cnt] = static_cast<long long>(t.offset); cnt++; }
    return cnt;
}
