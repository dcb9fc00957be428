"""GPU tier: the RSSI tap on port 0 (rssi_block.cpp:25-45) against the oracle's sequential restatement.  The CUDA kernel
evaluates window sums directly instead of the reference's running float sum, so the comparison is 1e-3 dB, for call sizes
below and above the 1024-sample horizon of the smoothing filter."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu


def test_rssi_tap_tracks_the_oracle(qrl, oracle):
    C, T = 3, 1 << 19
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=5600)
    X[1] *= 0.05                                               # a weak channel
    X[2, : T // 2] = 0                                         # silence first, then signal
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=200000)
    blk.enable_rssi(True)
    refs = [oracle.Rssi(0.0) for _ in range(C)]
    lo, i, sizes = 0, 0, [200000, 20000, 1000, 150000, 50, 33333]
    checked = 0
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        p0 = blk.read_port(0)
        got = blk.rssi(level=-3.5)
        for c in range(C):
            want = refs[c].work(p0[c]) - 3.5
            if lo > 150000:                                     # past the start-up of the 2000-sample window
                assert abs(float(got[c]) - want) < 1e-3 or want < -150.0, (lo, c, float(got[c]), want)
                checked += 1
    assert checked >= 9
    got = blk.rssi()
    assert got[0] - got[1] > 20.0                               # 26 dB weaker channel
