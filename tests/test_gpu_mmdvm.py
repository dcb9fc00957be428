"""GPU tier: the MMDVM multi-channel front end (gr_demod_mmdvm_multi2.cpp:30-127, gr_mod_mmdvm_multi2.cpp:28-131 without the MMDVM
protocol sink / source) through the C ABI against the CPU oracle: the per-channel chains either side of the filter bank (int16 and
IQ bit-identical, RSSI within float log10 tolerance), chunk invariance, the whole demodulator behind the polyphase channelizer, and
modulator -> demodulator loop-back on all channels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def lp2(qrl, gain, fs, fc, tw, att):
    import ctypes as C
    L = qrl.load_library()
    buf = np.zeros(16384, np.float32)
    n = L.qrl_firdes_low_pass_2(float(gain), float(fs), float(fc), float(tw), float(att), 5, buf.ctypes.data_as(C.c_void_p), len(buf))
    return buf[:n].copy()


def fm_rows(rng, C, n):
    t = np.arange(n)
    X = np.zeros((C, n), np.complex64)
    for c in range(C):
        a = 0.4 * np.sin(2 * np.pi * (400 + 130 * c) * t / 25000) + 0.2 * np.sin(2 * np.pi * 1700 * t / 25000 + c)
        ph = 2 * np.pi * 2500 * np.cumsum(a) / 25000
        amp = 0.05 * (1 + 0.8 * np.sin(2 * np.pi * 1.5 * t / 25000 + c))
        X[c] = (amp * np.exp(1j * ph) + 0.001 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    X[0, 3000:3700] = 0                                   # a silent stretch: atan2(0, 0), RSSI of zeros
    return X


@pytest.mark.parametrize("sizes", [(60000,), (1, 24, 25, 10007, 30000, 299, 19644)])
def test_rx_channel_chains_match_the_oracle(qrl, oracle, sizes):
    C = 4
    rng = np.random.default_rng(9400)
    X = fm_rows(rng, C, sum(sizes))
    rows = [2, 0, 5, 3]
    slab = np.zeros((6, X.shape[1]), np.complex64)
    for c, r in enumerate(rows):
        slab[r] = X[c]
    rx = qrl.MmdvmChannelsRx(C, rows=rows, n_rows=6, filter_width=5000, max_in=max(sizes))
    rx.calibrate_rssi(-12.5)
    outs, dbs, ats = [], [], []
    lo = 0
    for n in sizes:
        o, db, at = rx.work(slab[:, lo:lo + n]); lo += n
        outs.append(o); dbs.append(db); ats.append(at)
    got, gdb, gat = np.concatenate(outs, 1), np.concatenate(dbs, 1), np.concatenate(ats)
    for c in range(C):
        o = oracle.MmdvmRx(5000); o.calibrate_rssi(-12.5)
        want, wdb, wat = o.work(X[c])
        assert got.shape[1] == len(want) == X.shape[1] * 24 // 25
        assert np.array_equal(got[c], want), c
        assert np.array_equal(gat, wat) and len(wdb) == len(want) // 300
        assert np.max(np.abs(gdb[c] - wdb)) < 2e-4, c
    assert np.any(got != 0)


@pytest.mark.parametrize("sizes", [(48000,), (1, 23, 24, 7001, 20000, 20951)])
def test_tx_channel_chains_match_the_oracle(qrl, oracle, sizes):
    C = 3
    rng = np.random.default_rng(9500)
    n = sum(sizes)
    t = np.arange(n)
    S = np.stack([(6000 * np.sin(2 * np.pi * (500 + 200 * c) * t / 24000) + rng.integers(-300, 300, n)).astype(np.int16) for c in range(C)])
    S[1, :200] = 32767; S[2, 500:600] = -32768
    rows = [0, 9, 1]
    tx = qrl.MmdvmChannelsTx(C, rows=rows, n_rows=10, filter_width=5000, max_in=max(sizes))
    parts, lo = [], 0
    for m in sizes:
        parts.append(tx.work(S[:, lo:lo + m])); lo += m
    got = np.concatenate(parts, 1)
    assert got.shape == (10, n * 25 // 24)
    for c in range(C):
        want = oracle.MmdvmTx(5000).work(S[c])
        assert np.array_equal(got[rows[c]].view(np.uint32), want.view(np.uint32)), c
    for r in range(10):
        if r not in rows:
            assert not np.any(got[r])


def test_whole_demodulator_behind_the_channelizer(qrl, oracle):
    """250 ksps wideband with FM carriers on four of the ten 25 kHz slots -> channelizer -> port map -> int16: equal to the oracle's
    channelizer + per-channel chain; ragged wideband chunks."""
    rng = np.random.default_rng(9600)
    nch, n = 5, 250000
    t = np.arange(n)
    ports = qrl.mmdvm_port_map(nch)
    x = 0.002 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for c, p in enumerate(ports):
        a = 0.5 * np.sin(2 * np.pi * (300 + 170 * c) * t / 250000)
        ph = 2 * np.pi * 2000 * np.cumsum(a) / 250000
        f = p * 25000.0 if p < 5 else (p - 10) * 25000.0
        x = x + 0.08 * np.exp(1j * (ph + 2 * np.pi * f * t / 250000))
    x = x.astype(np.complex64)
    dem = qrl.MmdvmDemod(nch, 5000, max_in=100000)
    outs, lo = [], 0
    for m in (100000, 7, 33333, 99999, 16661):
        outs.append(dem.work(x[lo:lo + m])[0]); lo += m
    got = np.concatenate(outs, 1)
    taps = lp2(qrl, 1, 250000, 5000, 2000, 60)
    chan = oracle.PfbChannelizer(10, taps).work(x)
    for c, p in enumerate(ports):
        want = oracle.MmdvmRx(5000).work(chan[p])[0]
        m = min(got.shape[1], len(want))
        assert m > 20000 and abs(got.shape[1] - len(want)) <= 1
        assert np.array_equal(got[c, :m], want[:m]), c
        seg = got[c, 4000:m].astype(np.float64)              # the channel's tone came through the discriminator
        spec = np.abs(np.fft.rfft(seg * np.hanning(len(seg))))
        assert abs(np.argmax(spec[5:]) + 5 - (300 + 170 * c) * len(seg) / 24000) < 2, c


def test_modulator_to_demodulator_loop_back(qrl):
    nch, n = 7, 24000
    t = np.arange(n)
    S = np.stack([(4000 * np.sin(2 * np.pi * (350 + 90 * c) * t / 24000)).astype(np.int16) for c in range(nch)])
    mod = qrl.MmdvmMod(nch, 5000, max_in=n)
    mod.set_bb_gain(0.9)
    wide = mod.work(S)
    assert len(wide) == n * 25 // 24 * 10
    dem = qrl.MmdvmDemod(nch, 5000, max_in=len(wide))
    out = dem.work(wide)[0]
    for c in range(nch):
        seg = out[c, 3000:23000].astype(np.float64)
        tt = np.arange(len(seg))
        f = 350 + 90 * c
        A = np.stack([np.sin(2 * np.pi * f * tt / 24000), np.cos(2 * np.pi * f * tt / 24000), np.ones(len(seg))], 1)
        coef, *_ = np.linalg.lstsq(A, seg, rcond=None)
        assert abs(np.hypot(coef[0], coef[1]) - 4000) < 60, (c, coef)
        assert np.sqrt(np.mean((seg - A @ coef) ** 2)) < 120, c


def test_single_channel_mmdvm_blocks(qrl, oracle):
    """gr_mod_mmdvm / gr_demod_mmdvm (gr_mod_mmdvm.cpp:28-70, gr_demod_mmdvm.cpp:30-64; 250 ksps, x125/12 and x12/125, bb_gain in front of the
    resampler, RSSI tags in front of the channel filter, 10 kHz discriminator) for a batch of independent streams, ragged calls."""
    C, n = 3, 9000
    rng = np.random.default_rng(9700)
    t = np.arange(n)
    S = np.stack([(5000 * np.sin(2 * np.pi * (400 + 210 * c) * t / 24000) + rng.integers(-200, 200, n)).astype(np.int16) for c in range(C)])
    tx = qrl.MmdvmChannelsTx(C, filter_width=5000, max_in=5000, single=True)
    tx.set_bb_gain(0.7)
    parts, lo = [], 0
    for m in (1, 11, 12, 5000, 3976):
        parts.append(tx.work(S[:, lo:lo + m])); lo += m
    iq = np.concatenate(parts, 1)
    assert iq.shape == (C, n * 125 // 12)
    for c in range(C):
        o = oracle.MmdvmTx(5000, single=True); o.set_bb_gain(0.7)
        assert np.array_equal(iq[c].view(np.uint32), o.work(S[c]).view(np.uint32)), c
    rx = qrl.MmdvmChannelsRx(C, filter_width=5000, max_in=40000, single=True)
    rx.calibrate_rssi(3.0)
    outs, dbs, ats, lo = [], [], [], 0
    for m in (1, 124, 125, 40000, 20000, 33500):
        o, db, at = rx.work(iq[:, lo:lo + m]); lo += m
        outs.append(o); dbs.append(db); ats.append(at)
    assert lo == iq.shape[1]
    got, gdb, gat = np.concatenate(outs, 1), np.concatenate(dbs, 1), np.concatenate(ats)
    for c in range(C):
        o = oracle.MmdvmRx(5000, single=True); o.calibrate_rssi(3.0)
        want, wdb, wat = o.work(iq[c])
        assert got.shape[1] == len(want) == n and np.array_equal(got[c], want), c
        assert np.array_equal(gat, wat) and np.max(np.abs(gdb[c] - wdb)) < 2e-4, c


def zero_idle_case(qrl, oracle, single, q=4):
    """q = 4: the GPU tier's size; q = 1: a quarter of it (the emulated library in the CPU tier)"""
    C, n = 3, 3000 * q
    rng = np.random.default_rng(9800)
    t = np.arange(n)
    S = np.stack([(7000 * np.sin(2 * np.pi * (450 + 170 * c) * t / 24000) + rng.integers(-250, 250, n)).astype(np.int16) for c in range(C)])
    sizes = (1, 23, 24, 1000 * q, 750 * q, n - 48 - 1750 * q)
    # (call index before which the tag is registered, channel (-1: all), tagged int16 sample, count)
    tags = [(0, 0, 10, 75 * q), (0, 1, 100, 1250 * q), (0, 1, 500 * q, 10 * q), (3, -1, 1025 * q, 180 * q), (4, 2, 1725 * q, 2250 * q), (5, 0, 1250 * q, 750 * q)]
    tx = qrl.MmdvmChannelsTx(C, filter_width=5000, max_in=max(sizes), single=single)
    orc = [oracle.MmdvmTx(5000, single=single) for _ in range(C)]
    got, want, lo = [], [[] for _ in range(C)], 0
    for i, m in enumerate(sizes):
        for (at, ch, k, cnt) in tags:
            if at == i:
                tx.zero_samples(k, cnt, ch)
                for c in (range(C) if ch < 0 else [ch]):
                    orc[c].zero_samples(qrl.mmdvm_tag_item(k, single), cnt)
        got.append(tx.work(S[:, lo:lo + m]))
        for c in range(C):
            want[c].append(orc[c].work(S[c, lo:lo + m]))
        lo += m
    assert lo == n
    got = np.concatenate(got, 1)
    zeroed = 0
    for c in range(C):
        w = np.concatenate(want[c])
        assert got.shape[1] == len(w)
        assert np.array_equal(got[c].view(np.uint32), w.view(np.uint32)), c
        zeroed += int(np.sum(w == 0))
    assert zeroed > 1250 * q                              # the tags did clear stretches of the output
    # without tags the same handle type gives the plain chain (nothing is cleared by default)
    plain = qrl.MmdvmChannelsTx(C, filter_width=5000, max_in=n, single=single).work(S)
    assert np.sum(plain[0] == 0) < 40 and not np.array_equal(plain[0], got[0])


@pytest.mark.parametrize("single", [False, True])
def test_tx_zero_idle_bursts(qrl, oracle, single):
    """gr_zero_idle_bursts(0) on the MMDVM modulators (gr_mod_mmdvm.cpp:51-58 in front of the filter, gr_mod_mmdvm_multi2.cpp:88,108-117
    behind the x25/24 resampler): "zero_samples" tags per channel, a later tag overriding a running count, a count running across
    calls, a tag registered after its item went out -- bit-identical to the oracle's restatement of the same block, ragged calls."""
    zero_idle_case(qrl, oracle, single)
