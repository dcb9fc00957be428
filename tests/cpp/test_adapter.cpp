// C++ host-side test of include/qrl_b200_gr.hpp: the reference's factory surface + sink polling, end to end on the GPU.
// 4FSK TX -> 4FSK RX loop-back for 3 channels, bits delivered through gr_bit_sink::get_data() the way
// gr_modem::demodulate() polls it (gr_modem.cpp:1019-1117), then a bit-serial sync-word search (0xED89AA).
#include <cstdio>
#include <cstdlib>
#include <random>
#include "qrl_b200_gr.hpp"

using namespace qrl_gr;

int main()
{
    const int C = 3, NFR = 40;
    std::vector<std::vector<unsigned char>> payload(C);
    long n = 8 + NFR * 10 + 24;
    std::vector<unsigned char> bytes(static_cast<size_t>(C) * n, 0xAA);
    std::mt19937 rng(123);
    for (int c = 0; c < C; c++)
        for (int f = 0; f < NFR; f++) {
            unsigned char* p = &bytes[c * n + 8 + f * 10];
            p[0] = 0xED; p[1] = 0x89; p[2] = 0xAA;
            for (int k = 0; k < 7; k++) { p[3 + k] = rng() & 0xff; payload[c].push_back(p[3 + k]); }
        }
    try {
        auto mod = make_gr_mod_4fsk(25, 1000000, 1700, 3500, true, C, n);
        std::vector<gr_complex> iq; long n_out = 0;
        if (mod->work(bytes.data(), n, n, iq, &n_out) < 0) { std::printf("FAIL tx work\n"); return 1; }
        for (auto& v : iq) v *= 0.8f;
        auto demod = make_gr_demod_4fsk(5, 1000000, 1700, 3000, true, C, n_out);
        std::vector<gr_bit_sink> sinks(C);
        // feed in 3 chunks like a GNU Radio scheduler would
        long lo = 0; const long chunks[3] = { n_out / 3, n_out / 3 + 17, n_out - 2 * (n_out / 3) - 17 };
        std::vector<std::vector<unsigned char>> bits(C);
        for (int k = 0; k < 3; k++) {
            std::vector<gr_complex> slab(static_cast<size_t>(C) * chunks[k]);
            for (int c = 0; c < C; c++) std::copy(iq.begin() + c * n_out + lo, iq.begin() + c * n_out + lo + chunks[k], slab.begin() + c * chunks[k]);
            if (demod->work(slab.data(), static_cast<int>(chunks[k]), chunks[k]) < 0) { std::printf("FAIL rx work: %s\n", demod->last_error()); return 1; }
            lo += chunks[k];
            for (int c = 0; c < C; c++) {
                int nb = 0;
                const unsigned char* b = demod->port<unsigned char>(2, c, &nb);
                sinks[c].work(b, nb);
                if (auto* d = sinks[c].get_data()) { bits[c].insert(bits[c].end(), d->begin(), d->end()); delete d; }
            }
        }
        int total = 0;
        for (int c = 0; c < C; c++) {
            unsigned sh = 0; int found = 0;
            for (size_t i = 0; i + 56 < bits[c].size(); i++) {
                sh = ((sh << 1) | (bits[c][i] & 1)) & 0xffffff;
                if (sh == 0xED89AA) {
                    unsigned char fr[7];
                    for (int b = 0; b < 7; b++) { int t = 0; for (int k = 0; k < 8; k++) t = (t << 1) | (bits[c][i + 1 + b * 8 + k] & 1); fr[b] = t; }
                    for (int f = 0; f < NFR; f++) if (std::equal(fr, fr + 7, payload[c].begin() + f * 7)) { found++; break; }
                }
            }
            std::printf("channel %d: %d/%d frames\n", c, found, NFR);
            total += found;
        }
        if (total < C * (NFR - 4)) { std::printf("FAIL\n"); return 1; }
        // constructor failure surfaces as std::runtime_error
        bool threw = false;
        try { make_gr_demod_4fsk(7, 1000000, 1700, 3000, true); } catch (const std::runtime_error&) { threw = true; }
        if (!threw) { std::printf("FAIL: bad sps did not throw\n"); return 1; }
    } catch (const std::exception& e) { std::printf("FAIL exception: %s\n", e.what()); return 1; }
    std::printf("OK\n");
    return 0;
}
