"""GPU tier: AM receive chain (gr_demod_am.cpp: /50 -> complex band-pass -> squelch -> complex_to_mag -> agc2_ff ->
DC-blocking iir_filter_ffd -> x0.99 -> rational_resampler_fff(2,5) -> audio low-pass) against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def am_signal(rng, C, T):
    n = np.arange(T)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        aud = 0.5 * np.sin(2 * np.pi * (700.0 + 130.0 * c) * n / 1e6) + 0.3 * np.sin(2 * np.pi * 1900.0 * n / 1e6 + 0.4)
        x = 0.4 * (1.0 + 0.8 * aud) * np.exp(2j * np.pi * (rng.uniform(-300, 300) * n / 1e6 + rng.uniform(0, 1)))
        x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.004
        x[:int(rng.integers(0, 3000))] = 0                       # leading silence: the squelch gates it (variable rate)
        X[c] = x.astype(np.complex64)
    return X


@pytest.mark.parametrize("fw", [5000, 3000])
def test_am_parity_chunked(qrl, oracle, fw):
    C, T = 3, 400000
    X = am_signal(np.random.default_rng(60 + fw), C, T)
    blk = qrl.make_gr_demod_am(125, 1000000, 1700, fw, n_channels=C, max_samples=131072)
    acc = [[[] for _ in range(C)] for _ in range(2)]
    lo, i, sizes = 0, 0, [131072, 50, 33333, 1, 99999, 4096]
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(2):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_AM, 125, 1000000, 1700, fw, 0)
        rx.work(X[c])
        for p in range(2):
            got = np.concatenate(acc[p][c]); want = rx.port(p)
            assert len(got) == len(want) and len(want) > 1000, (p, len(got), len(want))
            assert np.array_equal(got, want), (c, p)
    audio = np.concatenate(acc[1][0])
    assert audio.dtype == np.float32 and float(np.std(audio[2000:])) > 0.05
