"""GPU tier: WBFM receive chain (gr_demod_wbfm.cpp: /5 -> 1363-tap low-pass at 200 ksps -> squelch -> quadrature demod -> x0.9 ->
de-emphasis IIR -> rational_resampler_fff(1,25)) against the CPU oracle; also exercises the shape-generic stage 1 with history."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def wbfm_signal(rng, C, T):
    n = np.arange(T)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        aud = 0.6 * np.sin(2 * np.pi * (900.0 + 170.0 * c) * n / 1e6) + 0.3 * np.sin(2 * np.pi * 3100.0 * n / 1e6 + 0.2)
        ph = 2 * np.pi * 50000.0 * np.cumsum(aud) / 1e6
        x = 0.5 * np.exp(1j * (ph + 2 * np.pi * rng.uniform(-500, 500) * n / 1e6))
        x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.005
        x[:int(rng.integers(0, 2000))] = 0
        X[c] = x.astype(np.complex64)
    return X


def test_wbfm_parity_chunked(qrl, oracle):
    C, T = 2, 300000
    X = wbfm_signal(np.random.default_rng(80), C, T)
    blk = qrl.make_gr_demod_wbfm(125, 1000000, 1700, 75000, n_channels=C, max_samples=131072)
    acc = [[[] for _ in range(C)] for _ in range(2)]
    lo, i, sizes = 0, 0, [131072, 7, 40001, 1, 99999]
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(2):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_WBFM, 125, 1000000, 1700, 75000, 0)
        rx.work(X[c])
        for p in range(2):
            got = np.concatenate(acc[p][c]); want = rx.port(p)
            assert len(got) == len(want) and len(want) > 1000, (p, len(got), len(want))
            assert np.array_equal(got, want), (c, p)
    audio = np.concatenate(acc[1][0])
    assert audio.dtype == np.float32 and float(np.std(audio[300:])) > 0.05
