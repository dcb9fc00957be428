"""GPU tier: SSB RX chain (/125 FIR 1045 taps, complex side-band filter, gated squelch, agc2, CESSB clipper +
stretcher, audio band-pass) against the CPU oracle, USB and LSB, streamed in uneven chunks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))


@pytest.mark.parametrize("sb", [0, 1])
def test_ssb_parity(qrl, oracle, sb):
    C, T = 3, 700000
    n = np.arange(T)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        rng = np.random.default_rng(40 + c)
        sign = -1.0 if sb else 1.0
        x = 0.05 * np.exp(2j * np.pi * sign * (700 + 150 * c) * n / 1e6) * (1 + 0.5 * np.sin(2 * np.pi * 3 * n / 1e6))
        x = x + 0.03 * np.exp(2j * np.pi * sign * 1900 * n / 1e6)
        x = x + 0.002 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))
        x[:1234] = 0
        X[c] = x.astype(np.complex64)
    blk = qrl.make_gr_demod_ssb(125, 1000000, 1700, 2700, sb, n_channels=C, max_samples=250000)
    acc = [[[] for _ in range(C)] for _ in range(2)]
    lo = 0; i = 0; sizes = [250000, 124, 99999, 1]
    while lo < T:
        m = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + m]); lo += m
        for p in range(2):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_SSB, 125, 1000000, 1700, 2700, sb)
        rx.work(X[c])
        w0, w1 = rx.port(0), rx.port(1)
        g0, g1 = np.concatenate(acc[0][c]), np.concatenate(acc[1][c])
        assert len(g0) == len(w0) and len(g1) == len(w1) and len(w1) > 5000, (len(g1), len(w1))
        assert rel_rms(g0, w0) <= 1e-5 and rel_rms(g1, w1) <= 1e-5
        assert np.array_equal(g0, w0) and np.array_equal(g1, w1)
        seg = g1[2000:5000].astype(np.float64)
        f = np.fft.rfftfreq(len(seg), 1 / 8000.0)
        assert abs(f[np.argmax(np.abs(np.fft.rfft(seg * np.hanning(len(seg)))))] - (700 + 150 * c)) < 10
