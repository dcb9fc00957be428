"""GPU tier: DSSS receive chain (gr_demod_dsss.cpp:32-124: /50 -> x13/50 -> Costas (snr) -> low-pass -> agc2 -> Barker-13 despreader ->
clock_recovery_mm_cc(omega = 1) -> Costas -> soft bits -> two Viterbi decoders -> descramblers) through the C ABI against the CPU oracle:
all four ports bit-identical, one call and ragged chunks.  The signal comes from the oracle's gr_mod_dsss restatement (8 bit/s: one
input byte is 10^6 samples)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dsss_channels(oracle, C, nbytes, seed):
    rng = np.random.default_rng(seed)
    X, datas = [], []
    for c in range(C):
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        iq = oracle.Tx(oracle.MOD_DSSS, 25, 1000000, 1700, 150, 0).work(data)
        n = np.arange(len(iq))
        x = iq * (0.4 + 0.1 * c) + 0.01 * (rng.standard_normal(len(iq)) + 1j * rng.standard_normal(len(iq)))
        x = x * np.exp(2j * np.pi * (0.4 - 0.5 * c) * n / 1e6 + 0.3j * (c + 1))
        X.append(x.astype(np.complex64)); datas.append(data)
    return np.stack(X), datas


def run_oracle(O, X):
    outs = []
    for c in range(X.shape[0]):
        rx = O.Rx(O.DEMOD_DSSS, 25, 1000000, 1700, 150, 0)
        rx.work(X[c])
        outs.append([rx.port(p) for p in range(4)])
    return outs


def test_dsss_parity_and_decoded_bytes(qrl, oracle):
    C, nbytes = 2, 26
    X, datas = dsss_channels(oracle, C, nbytes, 9100)
    want = run_oracle(oracle, X)
    T = 1 << 22
    blk = qrl.make_gr_demod_dsss(n_channels=C, max_samples=T)
    acc = [[[] for _ in range(C)] for _ in range(4)]
    for lo in range(0, X.shape[1], T):
        blk.work(X[:, lo:lo + T])
        for p in range(4):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        for p in range(4):
            got = np.concatenate(acc[p][c])
            assert got.dtype == want[c][p].dtype and len(got) == len(want[c][p]), (c, p, len(got), len(want[c][p]))
            assert np.array_equal(got, want[c][p]), (c, p)
        bits = np.unpackbits(datas[c])
        best = 0.0
        for port in (2, 3):
            g = np.concatenate(acc[port][c])
            for off in range(0, 40):
                m = min(len(g) - off, len(bits))
                if m > 60:          # the first symbols ride on the loops' acquisition: judge the settled part
                    best = max(best, float(np.mean(g[off + 48:off + m] == bits[48:m])))
        assert best == 1.0, c
    assert len(want[0][0]) == X.shape[1] // 50 * 13 // 50


def test_dsss_ragged_chunks(qrl, oracle):
    C, nbytes = 2, 14
    X, _ = dsss_channels(oracle, C, nbytes, 9200)
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_dsss(n_channels=C, max_samples=3000001)
    acc = [[[] for _ in range(C)] for _ in range(4)]
    sizes = [1, 49, 3000001, 123457, 2000000, 999, 1500000]
    lo, i = 0, 0
    while lo < X.shape[1]:
        n = min(sizes[i % len(sizes)], X.shape[1] - lo)
        blk.work(X[:, lo:lo + n]); lo += n; i += 1
        for p in range(4):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        for p in range(4):
            got = np.concatenate(acc[p][c])
            assert len(got) == len(want[c][p]) and np.array_equal(got, want[c][p]), (c, p)


@pytest.mark.parametrize("cuts", [(), (1, 3)])
def test_tx_dsss_matches_oracle_and_streams(qrl, oracle, cuts):
    """gr_mod_dsss (gr_mod_dsss.cpp:27-93, instance make_gr_mod_dsss(25, 1e6, 1700, 200)): bit-identical IQ, state carried across calls."""
    C, nbytes = 2, 4
    rng = np.random.default_rng(9300)
    data = rng.integers(0, 256, (C, nbytes), dtype=np.uint8)
    tx = qrl.make_gr_mod_dsss(n_channels=C, max_items=nbytes)
    tx.set_bb_gain(0.8)
    edges = [0, *cuts, nbytes]
    got = np.concatenate([tx.work(data[:, a:b]) for a, b in zip(edges[:-1], edges[1:])], axis=1)
    for c in range(C):
        o = oracle.Tx(oracle.MOD_DSSS, 25, 1000000, 1700, 200, 0)
        o.set_bb_gain(0.8)
        want = o.work(data[c])
        assert got.shape[1] == len(want) == nbytes * 1000000
        assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), c
