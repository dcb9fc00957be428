// stand-in for gnuradio/tags.h + the little of pmt the reference's gr_zero_idle_bursts.cpp uses (oracle/_ref build only)
#pragma once
#include <cstdint>
#include <memory>
#include <string>
namespace pmt {
struct pmt_base { std::string sym; uint64_t u64 = 0; float f32 = 0.0f; };
typedef std::shared_ptr<pmt_base> pmt_t;
inline pmt_t string_to_symbol(const std::string& s) { auto p = std::make_shared<pmt_base>(); p->sym = s; return p; }
inline pmt_t from_uint64(uint64_t v) { auto p = std::make_shared<pmt_base>(); p->u64 = v; return p; }
inline uint64_t to_uint64(const pmt_t& p) { return p->u64; }
inline pmt_t from_float(float v) { auto p = std::make_shared<pmt_base>(); p->f32 = v; return p; }
inline float to_float(const pmt_t& p) { return p->f32; }
inline bool eqv(const pmt_t& a, const pmt_t& b) { return a->sym == b->sym; }
}  // namespace pmt
namespace gr {
struct tag_t {
    uint64_t offset = 0;
    pmt::pmt_t key, value, srcid;
    static bool offset_compare(const tag_t& x, const tag_t& y) { return x.offset < y.offset; }
};
}  // namespace gr
