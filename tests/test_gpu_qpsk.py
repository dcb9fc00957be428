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
