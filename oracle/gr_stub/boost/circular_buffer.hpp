// stand-in for boost::circular_buffer (oracle/_ref build only; rx_fft_f, which the pinned tests do not drive, uses it)
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <class T> class circular_buffer {
public:
    void set_capacity(size_t c) { cap = c; if (v.size() > c) v.erase(v.begin(), v.begin() + (v.size() - c)); }
    void push_back(const T& x) { if (cap == 0) return; if (v.size() == cap) v.erase(v.begin()); v.push_back(x); }
    size_t size() const { return v.size(); }
    void clear() { v.clear(); }
    T* linearize() { return v.data(); }
private:
    std::vector<T> v; size_t cap = 0;
};
}
