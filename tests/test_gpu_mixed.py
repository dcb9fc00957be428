"""GPU tier: BASELINE config 4 in miniature -- a mixed FM / 4FSK / QPSK channel list is grouped by mode and
block-partitioned (qradiolink_b200.sharding), every (rank, mode) group is one batched handle, results match the
per-channel oracle.  Also exercises channel counts that are not multiples of 32."""
import numpy as np
import pytest

from qradiolink_b200 import sharding
from tests import siggen

pytestmark = pytest.mark.gpu


def test_mixed_modes_sharded_like_config4(qrl, oracle):
    T, NCH, WORLD = 1 << 18, 12, 2
    modes = [("nbfm", "4fsk", "qpsk")[ch % 3] for ch in range(NCH)]
    # per-channel signals (global channel index decides the seed)
    sig = {}
    for ch, m in enumerate(modes):
        if m == "nbfm":
            sig[ch] = siggen.gen_nbfm_channels(1, T, seed0=3000 + ch)[0]
        elif m == "4fsk":
            sig[ch] = siggen.gen_4fsk_channels(1, T, seed0=3000 + ch)[0][0]
        else:
            sig[ch] = siggen.gen_qpsk_channels(1, T, seed0=3000 + ch)[0][0]
    make = {"nbfm": lambda n: qrl.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=n, max_samples=T),
            "4fsk": lambda n: qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=n, max_samples=T),
            "qpsk": lambda n: qrl.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=n, max_samples=T)}
    okind = {"nbfm": (oracle.DEMOD_NBFM, 125, 1000000, 1700, 2500, 0), "4fsk": (oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1),
             "qpsk": (oracle.DEMOD_QPSK, 2, 1000000, 1700, 160000, 0)}
    seen = []
    for rank in range(WORLD):                   # the two "ranks" run one after the other on the single test GPU
        for m, chans in sharding.shard_channels(modes, WORLD, rank).items():
            blk = make[m](len(chans))
            blk.work(np.stack([sig[ch] for ch in chans]))
            last = blk.nports - 1
            out = blk.read_port(last)
            for i, ch in enumerate(chans):
                rx = oracle.Rx(*okind[m]); rx.work(sig[ch])
                want = rx.port(last)
                assert len(out[i]) == len(want) and np.array_equal(out[i], want), (rank, m, ch)
                seen.append(ch)
            blk.close()
    assert sorted(seen) == list(range(NCH))


@pytest.mark.parametrize("C", [1, 33, 65])
def test_channel_counts_not_multiple_of_32(qrl, oracle, C):
    T = 1 << 17
    base, _ = siggen.gen_4fsk_channels(3, T, seed0=4000)
    X = np.stack([np.roll(base[c % 3], 17 * c) for c in range(C)])
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    blk.work(X)
    bits = blk.read_port(2)
    for c in sorted({0, C // 2, C - 1}):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1); rx.work(X[c])
        assert np.array_equal(bits[c], rx.port(2)), c


@pytest.mark.parametrize("ntaps,D", [(419, 50), (100, 7), (33, 1), (5, 64)])
def test_standalone_decimating_fir_any_shape(qrl, oracle, ntaps, D):
    """qrl_fir_decim_ccf_device: shape-generic batched decimating FIR in THE FIR order, bit-identical to the oracle."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(ntaps + D)
    taps = (rng.standard_normal(ntaps) / ntaps).astype(np.float32)
    Cn, T = 3, 5003
    x = (rng.standard_normal((Cn, T)) + 1j * rng.standard_normal((Cn, T))).astype(np.complex64)
    nout = (T + D - 1) // D
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros((Cn, nout), dtype=torch.complex64, device="cuda")
    L = qrl.load_library()
    rc = L.qrl_fir_decim_ccf_device(taps.ctypes.data_as(C.c_void_p), ntaps, D, C.c_void_p(xd.data_ptr()), T, T,
                                    C.c_void_p(yd.data_ptr()), nout, Cn, None)
    assert rc == 0
    got = yd.cpu().numpy()
    for c in range(Cn):
        want = oracle.fir_decim_ccf(taps, D, x[c])
        assert len(want) == nout and np.array_equal(got[c], want)
