"""GPU tier: BPSK-2k and 2FSK-2k-FM RX chains (fll_band_edge_cc, clock_recovery_mm_cc / symbol_sync_ff, Costas(2),
two CCSDS decoders with the second behind delay(1)) against the CPU oracle: all four ports."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu


def make_signals(oracle, kind, C, T, seed):
    rng = np.random.default_rng(seed)
    X = np.zeros((C, T), np.complex64); payloads = []
    for c in range(C):
        data, pl = siggen.frames_4fsk(rng, max(4, int(T / 1e6 * 2000 / 80) - 6))
        if kind == "bpsk":
            iq = oracle.Tx(oracle.MOD_BPSK, 250, 1000000, 1700, 2800, 0).work(data)
        else:
            iq = oracle.Tx(oracle.MOD_2FSK, 25, 1000000, 1700, 4000, 1).work(data)
        # realistic receive level: the FLL's loop gain scales with signal power
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-150, 150), phase=rng.uniform(0, 6.28), delay=int(rng.integers(0, 300)),
                              snr_db=18.0, amp=0.1, total=T)
        payloads.append(pl)
    return X, payloads


def check(qrl, oracle, blk, okind, args, X, payloads, chunks):
    C, T = X.shape
    acc = [[[] for _ in range(C)] for _ in range(4)]
    lo = 0; i = 0
    while lo < T:
        n = min(chunks[i % len(chunks)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(4):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(okind, *args)
        rx.work(X[c])
        want = [rx.port(p) for p in range(4)]
        got = [np.concatenate(acc[p][c]) for p in range(4)]
        assert len(got[0]) == len(want[0]) and np.array_equal(got[0], want[0]), "port0"
        for p in (1, 2, 3):
            n = min(len(got[p]), len(want[p]))
            assert n > 0 and len(want[p]) - n <= 160, (p, len(got[p]), len(want[p]))
            assert np.array_equal(got[p][:n], want[p][:n]), "port %d channel %d" % (p, c)
        good = max(siggen.count_good_frames(got[p], 0xED89AA, 24, 7, payloads[c])[0] for p in (2, 3))
        assert good >= len(payloads[c]) - 6, (good, len(payloads[c]))


def test_2fsk_2k_fm_parity(qrl, oracle):
    C, T = 3, 1 << 20
    X, payloads = make_signals(oracle, "2fsk", C, T, 31)
    blk = qrl.make_gr_demod_2fsk(5, 1000000, 1700, 4000, True, n_channels=C, max_samples=400000)
    check(qrl, oracle, blk, oracle.DEMOD_2FSK, (5, 1000000, 1700, 4000, 1), X, payloads, [400000, 123457, 33])


def test_bpsk_2k_parity(qrl, oracle):
    C, T = 3, 1 << 20
    X, payloads = make_signals(oracle, "bpsk", C, T, 32)
    blk = qrl.make_gr_demod_bpsk(5, 1000000, 1700, 2400, n_channels=C, max_samples=400000)
    check(qrl, oracle, blk, oracle.DEMOD_BPSK, (5, 1000000, 1700, 2400, 0), X, payloads, [400000, 99999, 51])


def test_2fsk_2k_band_filter_variant(qrl, oracle):
    """2FSK2K (fm=false): upper/lower complex band filters, magnitude ratio, rail, symbol filter (gr_demod_2fsk.cpp:137-149)."""
    C, T = 2, 1 << 20
    rng = np.random.default_rng(35)
    X = np.zeros((C, T), np.complex64); payloads = []
    for c in range(C):
        data, pl = siggen.frames_4fsk(rng, 20)
        iq = oracle.Tx(oracle.MOD_2FSK, 25, 1000000, 1700, 4000, 0).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-60, 60), delay=int(rng.integers(0, 300)), snr_db=20.0, amp=0.1, total=T)
        payloads.append(pl)
    blk = qrl.make_gr_demod_2fsk(5, 1000000, 1700, 4000, False, n_channels=C, max_samples=400000)
    check(qrl, oracle, blk, oracle.DEMOD_2FSK, (5, 1000000, 1700, 4000, 0), X, payloads, [400000, 77777, 5])


@pytest.mark.parametrize("sps,fw,tx_sps", [(5, 4000, 25), (10, 2000, 50)])
def test_gmsk_demod_parity(qrl, oracle, sps, fw, tx_sps):
    """gr_demod_gmsk (GMSK2K / GMSK1K instances): the 2FSK chain without the FLL, low-pass symbol filter, its own clock loop
    constants -- all four ports against the oracle, ragged chunks."""
    C, T = 2, 1 << 18
    rng = np.random.default_rng(70 + sps)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        data, _ = siggen.frames_4fsk(rng, 6)
        iq = oracle.Tx(oracle.MOD_2FSK, tx_sps, 1000000, 1700, fw, 1).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-100, 100), phase=rng.uniform(0, 6.28), delay=int(rng.integers(0, 300)),
                              snr_db=20.0, amp=0.4, total=T)
    blk = qrl.make_gr_demod_gmsk(sps, 1000000, 1700, fw, n_channels=C, max_samples=100000)
    acc = [[[] for _ in range(C)] for _ in range(4)]
    lo, i, sizes = 0, 0, [100000, 49, 50021, 1, 77777]
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(4):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_GMSK, sps, 1000000, 1700, fw, 0)
        rx.work(X[c])
        for p in range(4):
            got = np.concatenate(acc[p][c]); want = rx.port(p)
            assert len(got) == len(want) and len(want) > 100, (p, len(got), len(want))
            assert np.array_equal(got, want), (c, p)
