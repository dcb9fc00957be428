"""GPU tier: DMR modulator (gr_mod_dmr.cpp:27-93: bytes -> dibits -> map -> RRC(0.2) x5 -> x0.66666666 -> frequency modulator ->
gr_zero_idle_bursts -> x0.9 -> bb gain -> x125 / 3 rational interpolator) through the C ABI against the CPU oracle: bit-identical IQ,
state and "zero_samples" tags carried across calls, then CUDA TX -> CUDA DMR RX."""
import numpy as np
import pytest

from tests import siggen

pytestmark = [pytest.mark.gpu]

TAGS = [(100, 720), (110, 100), (1, 50), (180, 333), (181, 1440), (299, 40)]


@pytest.mark.parametrize("cuts", [(), (100,), (1, 7, 150, 182)])
def test_tx_dmr_matches_oracle_and_streams(qrl, oracle, cuts):
    C, nbytes = 3, 300
    rng = np.random.default_rng(5500)
    data = rng.integers(0, 256, (C, nbytes), dtype=np.uint8)
    tx = qrl.make_gr_mod_dmr(n_channels=C, max_items=nbytes)
    for off, val in TAGS:
        tx.zero_samples(off, val, channel=-1 if off != 180 else 1)       # one tag on channel 1 only
    edges = [0, *cuts, nbytes]
    got = np.concatenate([tx.work(data[:, a:b]) for a, b in zip(edges[:-1], edges[1:])], axis=1)
    for c in range(C):
        o = oracle.Tx(oracle.MOD_DMR, 125, 1000000, 1700, 5000, 0)
        for off, val in TAGS:
            if off != 180 or c == 1:
                o.zero_samples(off, val)
        want = o.work(data[c])
        assert got.shape[1] == len(want) == (nbytes * 4 * 5 * 125 + 2) // 3, (got.shape, len(want))
        assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32)), c
    assert np.all(got[:, :int(1439 * 125 / 3) - 10] == 0) and np.any(got[:, 70000:] != 0)


def test_tx_dmr_late_tags_and_bb_gain(qrl, oracle):
    """A tag registered after part of its range has been produced clears the rest of its count; set_bb_gain (gr_mod_dmr.cpp:96-99)
    between calls, negative too (the cleared items become -0.0 on both sides)."""
    rng = np.random.default_rng(5501)
    data = rng.integers(0, 256, (2, 240), dtype=np.uint8)
    tx = qrl.make_gr_mod_dmr(n_channels=2, max_items=128)
    os_ = [oracle.Tx(oracle.MOD_DMR, 125, 1000000, 1700, 5000, 0) for _ in range(2)]
    got, want = [], [[], []]
    steps = [(0, 80, [], 1.0), (80, 160, [(78, 400), (90, 64)], -0.5), (160, 240, [(150, 3000)], 0.25)]
    for a, b, tags, gain in steps:
        for off, val in tags:
            tx.zero_samples(off, val)
            for o in os_:
                o.zero_samples(off, val)
        tx.set_bb_gain(gain)
        for o in os_:
            o.set_bb_gain(gain)
        got.append(tx.work(data[:, a:b]))
        for c in range(2):
            want[c].append(os_[c].work(data[c, a:b]))
    got = np.concatenate(got, axis=1)
    for c in range(2):
        assert np.array_equal(got[c].view(np.uint32), np.concatenate(want[c]).view(np.uint32)), c


def test_tx_dmr_loops_back_through_the_cuda_receiver(qrl):
    rng = np.random.default_rng(5502)
    data = rng.integers(0, 256, 400, dtype=np.uint8)
    iq = qrl.make_gr_mod_dmr(n_channels=1, max_items=len(data)).work(data[None, :])[0]
    x = siggen.channel(iq, rng, fo_hz=30, phase=0.2, delay=97, snr_db=30, amp=0.5, total=len(iq) + 30000)
    rx = qrl.make_gr_demod_dmr(n_channels=1, max_samples=len(x))
    rx.work(x[None, :])
    bits, tx_bits = rx.read_port(2)[0], np.unpackbits(data)
    best = 0.0
    for off in range(400, 900):
        n = min(len(bits) - off, len(tx_bits)) - 900        # the last 1439 items stay in the zero-idle delay line
        if n > 1000:
            best = max(best, float(np.mean(bits[off + 200:off + n] == tx_bits[200:n])))
    assert best == 1.0


def test_zero_samples_refused_on_other_modulators(qrl):
    tx = qrl.make_gr_mod_m17(n_channels=1, max_items=64)
    with pytest.raises(qrl.QrlError):
        tx.zero_samples(10, 100)
