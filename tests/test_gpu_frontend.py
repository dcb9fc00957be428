"""GPU tier: front end at device rates >= 2 Msps (gr_demod_base.cpp:1303-1362: rotator at the device rate + /N decimator with the 83-tap
Blackman-Harris low-pass at 2 Msps) against the oracle: ragged chunks, retunes in mid-stream, and feeding the 4FSK demodulator on the
device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("samp_rate", [2000000, 5000000, 10000000])
def test_frontend_matches_oracle(qrl, oracle, samp_rate):
    C, T = 3, 600000
    rng = np.random.default_rng(samp_rate // 1000000)
    n = np.arange(T)
    X = np.stack([(0.6 * np.exp(2j * np.pi * (30000.0 * (c + 1)) * n / samp_rate) + 0.05 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))).astype(np.complex64)
                  for c in range(C)])
    fe = qrl.Frontend(samp_rate, n_channels=C, max_in=200000)
    ofe = [oracle.Frontend(samp_rate) for _ in range(C)]
    assert ofe[0].ntaps == {2000000: 83, 5000000: 209, 10000000: 419}[samp_rate]
    lo, i, sizes = 0, 0, [200000, 1, 77777, 13, 123456]
    offsets = {1: 30000.0, 3: -12500.0}          # retune before chunks 1 and 3
    while lo < T:
        m = min(sizes[i % len(sizes)], T - lo)
        if i in offsets:
            fe.set_carrier_offset(offsets[i])
            for o in ofe:
                o.set_carrier_offset(offsets[i])
            if i == 3:
                fe.set_carrier_offset(5000.0, channel=1); ofe[1].set_carrier_offset(5000.0)
        got = fe.work(X[:, lo:lo + m])
        for c in range(C):
            want = ofe[c].work(X[c, lo:lo + m])
            assert got.shape[1] == len(want), (i, c, got.shape, len(want))
            assert np.array_equal(got[c], want), (i, c)
        lo += m; i += 1


def test_frontend_feeds_the_demodulator_on_device(qrl, oracle):
    from tests import siggen
    C, T = 2, 1 << 19
    X1, _ = siggen.gen_4fsk_channels(C, T, seed0=8900)
    # 2 Msps capture of the same signal shifted by +25 kHz: zero-stuff + crude interpolation is enough for a parity test
    X2 = np.zeros((C, 2 * T), np.complex64)
    X2[:, ::2] = X1; X2[:, 1::2] = X1
    X2 = (X2 * np.exp(2j * np.pi * 25000.0 * np.arange(2 * T) / 2e6)).astype(np.complex64)
    fe = qrl.Frontend(2000000, n_channels=C, max_in=2 * T)
    fe.set_carrier_offset(25000.0)
    rx = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T + 8)
    import ctypes as Ct
    import torch
    Xd = torch.from_numpy(X2).cuda()
    nout = fe.work_device(Xd.data_ptr(), 2 * T, 2 * T)
    ptr, stride, items = fe.out_device()
    assert items == nout == T
    fe._L.qrl_frontend_read(fe._h, None, 0) if False else None
    torch.cuda.synchronize()
    import qradiolink_b200 as q
    q.load_library().qrl_frontend_read(fe._h, np.zeros((C, 1), np.complex64).ctypes.data_as(Ct.c_void_p), 1)     # syncs the front end's stream
    rx.work_device(ptr, items, stride)
    bits = rx.read_port(2)
    for c in range(C):
        o = oracle.Frontend(2000000); o.set_carrier_offset(25000.0)
        y = o.work(X2[c])
        r = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        r.work(y)
        assert np.array_equal(bits[c], r.port(2)) and len(bits[c]) > 500
