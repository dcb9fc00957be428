// Minimal stand-in for the GNU Radio 3.10 runtime headers the reference's in-tree blocks include.
// TEST INFRASTRUCTURE (oracle/_ref build only): just enough of gr::block / gr::sync_block / io_signature /
// gr::thread for /root/reference/src/gr/{gr_4fsk_discriminator,gr_deframer_bb,gr_bit_sink,gr_audio_sink,
// gr_const_sink,dsss_*_impl,cessb/*_impl}.c* to compile UNMODIFIED where they lie.  The scheduler is not
// modelled: oracle/ref_blocks_shim.cpp calls work()/general_work() directly with the buffers, history and
// output multiples a GNU Radio scheduler would present.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <memory>
#include <mutex>
#include <condition_variable>
#include <string>
#include <vector>
#include <algorithm>
#include <time.h>
#include <sys/types.h>
#include <gnuradio/tags.h>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

namespace boost { struct mutex : std::mutex { typedef std::unique_lock<std::mutex> scoped_lock; }; }

namespace gr {
namespace thread {
using mutex = std::mutex;
using scoped_lock = std::unique_lock<std::mutex>;
using condition_variable = std::condition_variable;
}  // namespace thread

class io_signature {
public:
    typedef std::shared_ptr<io_signature> sptr;
    int min_streams, max_streams;
    std::vector<int> sizes;
    static sptr make(int mn, int mx, int size) { auto s = std::make_shared<io_signature>(); s->min_streams = mn; s->max_streams = mx; s->sizes.assign(1, size); return s; }
    static sptr makev(int mn, int mx, const std::vector<int>& v) { auto s = std::make_shared<io_signature>(); s->min_streams = mn; s->max_streams = mx; s->sizes = v; return s; }
};

class block {
public:
    block() {}
    block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : d_name(name), d_in(in), d_out(out) {}
    virtual ~block() {}
    virtual void forecast(int noutput_items, gr_vector_int& req) { for (auto& r : req) r = noutput_items + static_cast<int>(d_history) - 1; }
    virtual int general_work(int, gr_vector_int&, gr_vector_const_void_star&, gr_vector_void_star&) { return -1; }
    void set_history(unsigned h) { d_history = h; }
    unsigned history() const { return d_history; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    void set_alignment(int) {}
    void set_relative_rate(double r) { d_relative_rate = r; }
    double relative_rate() const { return d_relative_rate; }
    void consume_each(int n) { d_consumed += n; }
    long take_consumed() { const long c = d_consumed; d_consumed = 0; return c; }
    void set_thread_priority(int) {}
    const std::string& name() const { return d_name; }
    // stream tags: the shim places tags (absolute offsets) on input 0 and advances the item counters like the scheduler
    uint64_t nitems_written(unsigned) const { return d_nitems_written; }
    uint64_t nitems_read(unsigned) const { return d_nitems_read; }
    void stub_add_input_tag(const tag_t& t) { d_in_tags.push_back(t); }
    void add_item_tag(unsigned, uint64_t offset, const pmt::pmt_t& key, const pmt::pmt_t& value) { tag_t t; t.offset = offset; t.key = key; t.value = value; d_out_tags.push_back(t); }
    std::vector<tag_t>& stub_out_tags() { return d_out_tags; }
    void stub_advance(uint64_t nread, uint64_t nwritten) { d_nitems_read += nread; d_nitems_written += nwritten; }
    void get_tags_in_window(std::vector<tag_t>& v, unsigned, uint64_t rel_start, uint64_t rel_end, const pmt::pmt_t& key)
    {
        v.clear();
        for (const auto& t : d_in_tags)
            if (t.offset >= d_nitems_read + rel_start && t.offset < d_nitems_read + rel_end && pmt::eqv(t.key, key)) v.push_back(t);
    }
protected:
    std::vector<tag_t> d_in_tags, d_out_tags;
    uint64_t d_nitems_read = 0, d_nitems_written = 0;
    std::string d_name;
    io_signature::sptr d_in, d_out;
    unsigned d_history = 1;
    int d_output_multiple = 1;
    double d_relative_rate = 1.0;
    long d_consumed = 0;
};

class sync_block : public block {
public:
    sync_block() {}
    sync_block(const std::string& name, io_signature::sptr in, io_signature::sptr out) : block(name, in, out) {}
    virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
};
}  // namespace gr

namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T* p) { return std::shared_ptr<T>(p); }
}
