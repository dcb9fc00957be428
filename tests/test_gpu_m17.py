"""GPU tier: M17 receive chain (gr_demod_m17.cpp: x3/125 rational resampler to 24 ksps, 4FSK at 5 samples per symbol, hard-decision
bit tail) against the CPU oracle: all three ports, ragged chunks.  Exercises the shape-generic rational stage 1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def m17_like_signal(rng, T):
    nsym = int(T / 1e6 * 4800) + 2
    sy = np.array([-1.5, -0.5, 0.5, 1.5])[rng.integers(0, 4, nsym)]
    t = np.arange(T) / 1e6
    x = sy[np.minimum((t * 4800).astype(int), nsym - 1)]
    k = np.hanning(400); k /= k.sum()
    xs = np.convolve(x, k, mode="same")
    ph = 2 * np.pi * np.cumsum(xs * 800.0) / 1e6
    iq = 0.5 * np.exp(1j * (ph + 2 * np.pi * rng.uniform(-100, 100) * np.arange(T) / 1e6))
    iq = iq + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.01
    return iq.astype(np.complex64)


def test_m17_parity_chunked(qrl, oracle):
    C, T = 3, 400000
    rng = np.random.default_rng(90)
    X = np.stack([m17_like_signal(rng, T) for _ in range(C)])
    blk = qrl.make_gr_demod_m17(n_channels=C, max_samples=150000)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    lo, i, sizes = 0, 0, [150000, 41, 125, 66667, 1, 99991]
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_M17, 125, 1000000, 1700, 9000, 0)
        rx.work(X[c])
        for p in range(3):
            got = np.concatenate(acc[p][c]); want = rx.port(p)
            assert len(got) == len(want) and len(want) > 1000, (p, len(got), len(want))
            assert np.array_equal(got, want), (c, p)
    bits = np.concatenate(acc[2][0])
    assert bits.max() == 1 and 0.05 < bits.mean() < 0.95
