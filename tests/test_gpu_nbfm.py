"""GPU tier: CUDA NBFM RX chain (/50 FIR, channel LPF, gated power squelch, quadrature demod, 2/5 resampler, audio
LPF, double-precision de-emphasis IIR) against the CPU oracle: float ports within 1e-5 RMS, identical counts."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))


@pytest.mark.parametrize("fw", [2500, 5000])
def test_nbfm_parity(qrl, oracle, fw):
    C, T = 3, 1 << 19
    X = siggen.gen_nbfm_channels(C, T, seed0=1, lead_zeros=777)
    blk = qrl.make_gr_demod_nbfm(125, 1000000, 1700, fw, n_channels=C, max_samples=T)
    blk.work(X)
    p0, p1 = blk.read_port(0), blk.read_port(1)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_NBFM, 125, 1000000, 1700, fw, 0)
        rx.work(X[c])
        w0, w1 = rx.port(0), rx.port(1)
        assert len(p0[c]) == len(w0) and len(p1[c]) == len(w1) and len(w1) > 3000, (len(p1[c]), len(w1))
        assert rel_rms(p0[c], w0) <= 1e-5 and rel_rms(p1[c], w1) <= 1e-5
        assert np.array_equal(p1[c], w1)            # same operation order: bit-identical in practice
        # the audio is the two tones: spectral peak at the stronger one
        seg = p1[c][2000:4000].astype(np.float64)
        f = np.fft.rfftfreq(len(seg), 1 / 8000.0)
        pk = f[np.argmax(np.abs(np.fft.rfft(seg * np.hanning(len(seg)))))]
        assert abs(pk - (1000.0 + 37.0 * c)) < 12.0


def test_nbfm_chunked_and_sinks(qrl, oracle):
    C, T = 2, 260000
    X = siggen.gen_nbfm_channels(C, T, seed0=9)
    blk = qrl.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=C, max_samples=50000)
    sinks = [qrl.gr_audio_sink() for _ in range(C)]
    packets = [[] for _ in range(C)]
    acc = [[] for _ in range(C)]
    lo = 0
    sizes = [1, 49, 50000, 12345, 50000, 33333]
    i = 0
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for c, v in enumerate(blk.read_port(1)):
            acc[c].append(v)
            sinks[c].work(v)
            while True:
                d = sinks[c].get_data()         # 640-sample packets, like gr_modem::demodulateAnalog polls them
                if d is None:
                    break
                packets[c].append(d)
    assert lo == T
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_NBFM, 125, 1000000, 1700, 2500, 0)
        rx.work(X[c])
        w1 = rx.port(1)
        g1 = np.concatenate(acc[c])
        assert len(g1) == len(w1) and np.array_equal(g1, w1)
        assert len(packets[c]) == len(w1) // 640 and all(len(p) == 640 for p in packets[c])
