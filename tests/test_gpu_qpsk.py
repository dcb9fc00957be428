"""GPU tier: CUDA QPSK-250k RX chain (/2 FIR + RRC + agc2 + Costas PLL + symbol sync + Costas + diff phasor +
CCSDS Viterbi + descrambler) against the CPU oracle.  Bits bit-exact, float ports <= 1e-5 RMS."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))


def run_oracle(O, X):
    outs = []
    for c in range(X.shape[0]):
        rx = O.Rx(O.DEMOD_QPSK, 2, 1000000, 1700, 160000, 0)
        rx.work(X[c])
        outs.append((rx.port(0), rx.port(1), rx.port(2)))
    return outs


def test_qpsk_parity_and_frames(qrl, oracle):
    C, T = 3, 1 << 18
    X, payloads = siggen.gen_qpsk_channels(C, T, seed0=2000)
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
    blk.work(X)
    got = [blk.read_port(p) for p in range(3)]
    for c in range(C):
        w0, w1, w2 = want[c]
        assert len(got[0][c]) == len(w0) and len(got[1][c]) == len(w1) and len(got[2][c]) == len(w2), \
            (len(got[0][c]), len(w0), len(got[1][c]), len(w1), len(got[2][c]), len(w2))
        assert np.array_equal(got[2][c], w2), "decoded bits differ on channel %d" % c
        assert rel_rms(got[0][c], w0) <= 1e-5 and rel_rms(got[1][c], w1) <= 1e-5
        assert np.array_equal(got[0][c], w0) and np.array_equal(got[1][c], w1)
        good, found = siggen.count_good_frames(got[2][c], 0xDE98AA, 24, 1516, payloads[c])
        assert good == len(payloads[c]) and good >= 3


def test_qpsk_chunked_stream(qrl, oracle):
    C, T = 2, 300000
    X, _ = siggen.gen_qpsk_channels(C, T, seed0=2100)
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=100001)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    sizes = [1, 2, 3, 100001, 7777, 65536, 99]
    lo = 0; i = 0
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo)
        blk.work(X[:, lo:lo + n]); lo += n; i += 1
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        w0, w1, w2 = want[c]
        g0, g1, g2 = (np.concatenate(acc[p][c]) for p in range(3))
        assert len(g0) == len(w0) and np.array_equal(g0, w0)
        n1 = min(len(g1), len(w1)); n2 = min(len(g2), len(w2))
        assert n1 >= len(w1) - 8 and n2 >= len(w2) - 160
        assert np.array_equal(g1[:n1], w1[:n1]) and np.array_equal(g2[:n2], w2[:n2])


@pytest.mark.parametrize("tx_sps,rx_sps,fw", [(100, 25, 6500), (500, 125, 1300)])
def test_qpsk_fll_variants(qrl, oracle, tx_sps, rx_sps, fw):
    """QPSK20K (make_gr_demod_qpsk(25,...,6500): /25, 681 taps) and QPSK2K (125,...,1300: /100, 2727 taps): both put a
    fll_band_edge_cc between the resampler and the shaping filter (gr_demod_qpsk.cpp:130-134)."""
    C, T = 2, 1 << 20
    rng = np.random.default_rng(61)
    X = np.zeros((C, T), np.complex64); payloads = []
    for c in range(C):
        data, pl = siggen.frames_4fsk(rng, 30 if tx_sps == 500 else 150)
        iq = oracle.Tx(oracle.MOD_QPSK, tx_sps, 1000000, 1700, fw, 0).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-40, 40), phase=rng.uniform(0, 6.28), delay=int(rng.integers(0, 200)),
                              snr_db=18.0, amp=0.2, total=T)
        payloads.append(pl)
    blk = qrl.make_gr_demod_qpsk(rx_sps, 1000000, 1700, fw, n_channels=C, max_samples=600000)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    for lo, hi in ((0, 600000), (600000, 600001), (600001, T)):
        blk.work(X[:, lo:hi])
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_QPSK, rx_sps, 1000000, 1700, fw, 0)
        rx.work(X[c])
        for p in range(3):
            got, want = np.concatenate(acc[p][c]), rx.port(p)
            n = min(len(got), len(want))
            assert n > 0 and len(want) - n <= 160, (p, len(got), len(want))
            assert np.array_equal(got[:n], want[:n]), (c, p)
        good, found = siggen.count_good_frames(np.concatenate(acc[2][c]), 0xED89AA, 24, 7, payloads[c])
        assert good >= 10


def test_qpsk_hot_and_fading_input(qrl, oracle):
    """Channel levels above the AGC reference (gain < 1) with 30 dB dips: the region where the complex AGC's signed rate compare
    matters (attack 1.0 / decay 0.1, gr_demod_qpsk.cpp:97).  Ports identical to the oracle, frames still recovered."""
    C, T = 3, 1 << 18
    X, payloads = siggen.gen_qpsk_channels(C, T, seed0=2400)
    n = np.arange(T)
    for c in range(C):
        level = np.where((n // 37000) % 3 == 2, 0.12, 5.0 + 2.0 * c).astype(np.float32)
        X[c] *= level
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
    blk.work(X)
    got = [blk.read_port(p) for p in range(3)]
    for c in range(C):
        for p in range(3):
            assert len(got[p][c]) == len(want[c][p]) and np.array_equal(got[p][c], want[c][p]), (c, p)
