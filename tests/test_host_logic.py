"""CPU tier: host-side logic and exact-equivalence arguments the CUDA kernels rely on."""
import numpy as np


def test_rect4_slicer_threshold_form_is_exact():
    """kernels use re >= {-1, -2^-24, 1-2^-23} instead of clamp(floor(re + 2.0f), 0, 3) (constellation_rect
    sector search): check equality on random floats and on every float near the three boundaries."""
    rng = np.random.default_rng(0)
    xs = [rng.standard_normal(2_000_000).astype(np.float32) * 2,
          rng.uniform(-3, 3, 2_000_000).astype(np.float32)]
    for centre in (-1.0, 0.0, 1.0, -2.0, 2.0):
        v = np.float32(centre)
        lo = [v]; hi = [v]
        for _ in range(4096):
            lo.append(np.nextafter(lo[-1], np.float32(-np.inf), dtype=np.float32))
            hi.append(np.nextafter(hi[-1], np.float32(np.inf), dtype=np.float32))
        xs.append(np.array(lo + hi, np.float32))
    xs.append(np.array([0.0, -0.0, 1e-38, -1e-38, 1e-45, -1e-45, -5.9604645e-8, -5.9604652e-8, 0.99999994, 0.9999999], np.float32))
    x = np.concatenate(xs)
    sec = np.clip(np.floor((x + np.float32(2.0)).astype(np.float32)), 0, 3).astype(np.int32)
    want = (np.float32(-1.5) + sec.astype(np.float32)).astype(np.float32)
    t3, t2, t1 = np.float32(0.99999988079071044921875), np.float32(-5.9604644775390625e-8), np.float32(-1.0)
    got = np.where(x >= t3, 1.5, np.where(x >= t2, 0.5, np.where(x >= t1, -0.5, -1.5))).astype(np.float32)
    assert np.array_equal(got, want)


def test_viterbi_metric_spread_bound():
    """32-bit metrics without renormalisation decide like VOLK's 8-bit kernel iff the 8-bit one never wraps:
    worst case spread over K-1 = 6 steps of branch cost <= 31 plus one add stays below 256."""
    assert 6 * 31 + 31 < 256 and 63 + 5 * 31 + 31 < 256


def test_framing_table_and_port_map_match_the_reference_literals():
    """MODE_FRAMING restates gr_modem::toggleRxMode (gr_modem.cpp:203-322: _bit_buf_len, _rx_frame_length) and the findSync branch
    each modem type takes (gr_modem.cpp:1208-1274); mmdvm_port_map restates the channel -> port wiring of
    gr_demod_mmdvm_multi2.cpp:110-124."""
    from qradiolink_b200 import framing, pfb
    F = framing.MODE_FRAMING
    assert F["4FSK2KFM"] == (framing.SYNC_NARROW, 8 * 8, 7) and F["4FSK2K"] == F["4FSK2KFM"] == F["BPSK2K"] == F["QPSK2K"]
    assert F["QPSK250K"] == (framing.SYNC_WIDE, 1517 * 8, 1516) and F["4FSK100K"] == (framing.SYNC_WIDE, 623 * 8, 622)
    assert F["QPSKVideo"] == (framing.SYNC_WIDE, 3123 * 8, 3122)
    assert F["QPSK20K"] == F["4FSK10KFM"] == F["2FSK10KFM"] == (framing.SYNC_NARROW, 48 * 8, 47)
    for m in ("BPSK1K", "2FSK1KFM", "2FSK1K", "4FSK1KFM"):
        assert F[m] == (framing.SYNC_1K, 4 * 8, 4)
    assert (framing.FrameTypeVoice, framing.FrameTypeVoice1, framing.FrameTypeEnd) == (0xED89, 0xB5, 0x4C8A2B)
    assert pfb.mmdvm_port_map(3) == [0, 1, 2] and pfb.mmdvm_port_map(7) == [0, 1, 2, 3, 9, 8, 7]


def test_costas_phase_wrap_in_float_is_exact():
    """qrl_costas4_snr_chunk wraps the loop phase without double precision: (a - C) + float(C - 2 pi) must equal the oracle's
    (float)((double)a - 2 pi) for EVERY float a the loop can hold when it wraps (|a| in [C, C + 1 + alpha], C = 6.28318548f)."""
    C = np.float32(6.2831854820251465)
    twopi = 2.0 * 3.14159265358979323846
    n = 3_400_000                                             # all floats from C up to 7.90 (> C + 1.6)
    a = (np.uint32(C.view(np.uint32)) + np.arange(n, dtype=np.uint32)).view(np.float32)
    assert a[0] == C and a[-1] > C + np.float32(1.6)
    want = (a.astype(np.float64) - twopi).astype(np.float32)
    delta = np.float32(float(C) - twopi)
    assert delta == np.float32(1.7484555314695172e-07)
    got = (a - C) + delta
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the negative side is the mirror image (IEEE rounding is sign-symmetric)
    want_n = ((-a).astype(np.float64) + twopi).astype(np.float32)
    assert np.array_equal((-got).view(np.uint32), want_n.view(np.uint32))


def test_mmdvm_tag_item_is_the_rounded_rate_scaling():
    """qradiolink_b200.mmdvm_tag_item: a "zero_samples" tag on int16 sample k reaches gr_zero_idle_bursts(0) on item k for gr_mod_mmdvm
    (1:1 blocks in front of it) and on floor(k * 25 / 24 + 1/2) behind the x25/24 resampler of gr_mod_mmdvm_multi2 (exact rational
    arithmetic, as GNU Radio's scheduler scales tag offsets across a rate-changing block)."""
    from fractions import Fraction
    from math import floor
    from qradiolink_b200.mmdvm import mmdvm_tag_item
    for k in list(range(0, 200)) + [719, 720, 721, 1439, 24000 * 3600 + 11, 2 ** 40 + 5]:
        assert mmdvm_tag_item(k, single=True) == k
        assert mmdvm_tag_item(k, single=False) == floor(Fraction(k) * Fraction(25, 24) + Fraction(1, 2))
    # monotone, and every 24 input samples advance the item by exactly 25
    assert all(mmdvm_tag_item(k + 24) - mmdvm_tag_item(k) == 25 for k in range(100))


def test_costas_tanh_table_selects_in_front_of_the_load_are_exact():
    """qrl_costas4_snr_chunk reads tanhf_lut(x) as ONE table load: index floor(fma(64, x, 128)) for -2 < x <= 2, entry 257 (= 1.0f) for
    x > 2, entry 258 (= -1.0f) for x <= -2 or NaN, entry 256 = entry 255 for x = 2 exactly.  Must equal the oracle's tanhf_lut
    (oracle/qrl_oracle.c: two compares, (int)(128.0f + 64.0f * x), index clamped to 255) for every float: checked over ALL floats of
    [-2.5, -1.5], [1.5, 2.5] and around the 256 step positions k / 64, plus a dense sweep and the special values."""
    import ctypes as C
    from oracle import oracle as O
    tab = np.zeros(256, np.float32)
    O.lib().qo_tanh_table.argtypes = [C.c_void_p]
    O.lib().qo_tanh_table(tab.ctypes.data_as(C.c_void_p))
    ext = np.concatenate([tab, [tab[255], np.float32(1.0), np.float32(-1.0)]]).astype(np.float32)      # qrl_fill_tanh_s

    def oracle_lut(x):
        with np.errstate(over="ignore", invalid="ignore"):      # +-3e38, inf: the branches in front of the cast take them, as in C
            idx = np.minimum(np.nan_to_num(np.float32(128.0) + np.float32(64.0) * x, nan=0.0, posinf=1e9, neginf=-1e9).astype(np.int64), 255)
        idx = np.clip(idx, 0, 255)
        return np.where(x > 2, np.float32(1.0), np.where(x <= -2, np.float32(-1.0), tab[idx])).astype(np.float32)

    def kernel_lut(x):
        with np.errstate(over="ignore", invalid="ignore"):
            v = np.float32(64.0) * x + np.float32(128.0)         # 64 x is exact in float32, so mul + add rounds once: the kernel's fma
            sel = np.where(x > 2, np.float32(257.0), np.where(x > -2, v, np.float32(258.0))).astype(np.float32)
            idx = np.floor(sel).astype(np.int64)                     # FADD.RM onto 2^23: floor
        assert idx.min() >= 0 and idx.max() <= 258               # the load address never leaves the table, NaN included
        return ext[idx]

    def all_floats(lo, hi):
        a, b = np.float32(lo).view(np.uint32), np.float32(hi).view(np.uint32)
        lo_u, hi_u = (a, b) if a <= b else (b, a)
        return np.arange(lo_u, hi_u + 1, dtype=np.uint32).view(np.float32)

    chunks = [all_floats(1.5, 2.5), all_floats(-1.5, -2.5), np.linspace(-3, 3, 2_000_001, dtype=np.float32)]
    for k in range(-128, 129):                                    # 2048 floats either side of every table step
        c = np.float32(k / 64.0)
        u = c.view(np.uint32).astype(np.int64)
        if k == 0:
            chunks.append(np.array([0.0, -0.0, 1e-45, -1e-45, 1e-30, -1e-30], np.float32))
        else:
            chunks.append(np.arange(u - 2048, u + 2049).astype(np.uint32).view(np.float32))
    x = np.concatenate(chunks)
    assert np.array_equal(kernel_lut(x).view(np.uint32), oracle_lut(x).view(np.uint32))
    special = np.array([np.inf, -np.inf, 3e38, -3e38, 2.0, -2.0, np.nextafter(np.float32(2), np.float32(3)), np.nextafter(np.float32(-2), np.float32(0))], np.float32)
    assert np.array_equal(kernel_lut(special).view(np.uint32), oracle_lut(special).view(np.uint32))
    assert kernel_lut(np.array([np.nan], np.float32))[0] == -1.0   # defined (entry 258); the reference's (int)NaN is not


def test_symbol_sync_uniform_trip_count_is_safe():
    """symsync_kernel's lean loops let every lane take ksafe = floor(max(left - 1, 0) * 0.9999 / pmax) + 1 symbols without looking at the
    window end (left = rows between the lane's position and the last allowed start, pmax = max_period + |alpha|).  Worst case in float32
    arithmetic as the kernel does it (ph = mu + inst, floor by two compares): every instantaneous period at pmax, the fractional phase
    starting just below 1 -- no symbol of the batch may start beyond the limit, for the loop constants of every lean instance."""
    for sps, dev, alpha in ((10.0, 0.05, 0.0439), (2.0, 0.0008, 2.6e-4), (4.0, 0.002, 1.0e-3), (5.0, 0.05, 0.03), (12.0, 0.05, 0.0439)):
        pmax = np.float32(sps + dev) + np.float32(abs(alpha))
        inv_s = np.float32(0.9999) / pmax
        for left in list(range(0, 40)) + [95, 96, 127, 255, 479, 480, 481, 511, 1000]:
            ksafe = int(np.float32(max(left - 1.0, 0.0)) * inv_s) + 1
            o, mu = 0, np.nextafter(np.float32(1.0), np.float32(0.0))
            for j in range(ksafe):
                assert o <= left, (sps, left, ksafe, j, o)        # symbol j starts at row o: must not be beyond the last allowed start
                ph = np.float32(mu + pmax)
                fl = np.floor(ph)
                mu = np.float32(ph - fl)
                o += int(fl)
            # and the bound is tight enough to be useful: at most two symbols short of what the window holds
            assert ksafe >= int(left / float(pmax)) - 1
