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
