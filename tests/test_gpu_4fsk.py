"""GPU tier: the CUDA 4FSK-2k-FM RX chain (through the C ABI) against the CPU oracle on the same seeded input.
Integer ports must be bit-exact; float ports within 1e-5 RMS (north_star tolerance)."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))


def run_oracle(O, X, fw=3000):
    outs = []
    for c in range(X.shape[0]):
        rx = O.Rx(O.DEMOD_4FSK, 5, 1000000, 1700, fw, 1)
        rx.work(X[c])
        outs.append((rx.port(0), rx.port(1), rx.port(2)))
    return outs


def compare(got_ports, want, exact_float=False):
    p0, p1, p2 = got_ports
    for c, (w0, w1, w2) in enumerate(want):
        assert len(p0[c]) == len(w0) and len(p1[c]) == len(w1) and len(p2[c]) == len(w2), (c, len(p1[c]), len(w1), len(p2[c]), len(w2))
        assert np.array_equal(p2[c], w2), "decoded bits differ on channel %d" % c
        assert rel_rms(p0[c], w0) <= 1e-5 and rel_rms(p1[c], w1) <= 1e-5
        if exact_float:
            assert np.array_equal(p0[c], w0) and np.array_equal(p1[c], w1)


def test_4fsk_fm_parity_single_call(qrl, oracle):
    C, T = 5, 1 << 19
    X, payloads = siggen.gen_4fsk_channels(C, T, seed0=1000)
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    blk.work(X)
    got = [blk.read_port(p) for p in range(3)]
    compare(got, want, exact_float=True)
    good, found = siggen.count_good_frames(got[2][0], 0xED89AA, 24, 7, payloads[0])
    assert good >= 3
    assert blk.launches >= 5


def test_4fsk_fm_chunk_invariance_and_ragged(qrl, oracle):
    """Feed the same stream in odd-sized chunks (not multiples of the decimation): state must carry."""
    C, T = 3, 400000
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=1100)
    want = run_oracle(oracle, X)
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    sizes = [1, 49, 50, 51, 70001, 12345, 33333, 7]
    lo = 0; i = 0
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo)
        blk.work(X[:, lo:lo + n]); lo += n; i += 1
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    got = [[np.concatenate(acc[p][c]) for c in range(C)] for p in range(3)]
    # the oracle fed in one piece produces the same stream; only the tail latency may differ
    for c in range(C):
        w0, w1, w2 = want[c]
        assert len(got[0][c]) == len(w0) and np.array_equal(got[0][c], w0)
        n1 = min(len(got[1][c]), len(w1)); n2 = min(len(got[2][c]), len(w2))
        assert n1 >= len(w1) - 2 and n2 >= len(w2) - 80
        assert np.array_equal(got[1][c][:n1], w1[:n1]) and np.array_equal(got[2][c][:n2], w2[:n2])


def test_4fsk_empty_and_silence(qrl, oracle):
    C, T = 2, 100000
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    X = np.zeros((C, T), np.complex64)
    blk.work(X)
    want = run_oracle(oracle, X)
    got = [blk.read_port(p) for p in range(3)]
    compare(got, want, exact_float=True)
    blk.work(np.zeros((C, 0), np.complex64))          # empty chunk is a no-op
    with pytest.raises(qrl.QrlError):
        blk.work(np.zeros((C, T + 1), np.complex64))  # larger than max_samples -> QRL_ERANGE


def test_full_size_properties(qrl, oracle):
    """BASELINE config 2 shape (64 ch) at a size the oracle cannot sweep quickly: check size-independent
    properties -- every channel recovers its own frames, and channels do not leak into each other."""
    C, T = 64, 1 << 20
    X, payloads = siggen.gen_4fsk_channels(C, T, seed0=1000)
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    blk.work(X)
    bits = blk.read_port(2)
    for c in range(C):
        good, found = siggen.count_good_frames(bits[c], 0xED89AA, 24, 7, payloads[c])
        assert good >= len(payloads[c]) - 4 and found - good <= 1, (c, good, found)
    # spot-check 3 channels bit-exact against the oracle
    for c in (0, 31, 63):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        rx.work(X[c])
        assert np.array_equal(bits[c], rx.port(2))


def test_4fsk_1k_fm_parity_d100(qrl, oracle):
    """4FSK1KFM: make_gr_demod_4fsk(10,...,2000,true) -> /100 decimator with 837 taps (gr_demod_base.cpp:213)."""
    C, T = 2, 1 << 19
    rng = np.random.default_rng(77)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        data = rng.integers(0, 256, 120, dtype=np.uint8)
        iq = oracle.Tx(oracle.MOD_4FSK, 50, 1000000, 1700, 2000, 1).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-100, 100), delay=int(rng.integers(0, 300)), snr_db=20.0, total=T)
    blk = qrl.make_gr_demod_4fsk(10, 1000000, 1700, 2000, True, n_channels=C, max_samples=T)
    blk.work(X[:, :200001]); a = [blk.read_port(p) for p in range(3)]
    blk.work(X[:, 200001:]); b = [blk.read_port(p) for p in range(3)]
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 10, 1000000, 1700, 2000, 1)
        rx.work(X[c])
        for p in range(3):
            got = np.concatenate([a[p][c], b[p][c]])
            want = rx.port(p)
            n = min(len(got), len(want))
            assert n > 0 and len(want) - n <= 80
            assert np.array_equal(got[:n], want[:n]), (c, p)


def test_4fsk_2k_discriminator_variant(qrl, oracle):
    """4FSK2K (fm=false): four complex band-pass filters + gr_4fsk_discriminator + 837-tap symbol filter + symbol_sync_cc."""
    C, T = 2, 1 << 20
    X, payloads = siggen.gen_4fsk_channels(C, T, seed0=1300, fm=False)
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 4000, False, n_channels=C, max_samples=600000)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    for lo, hi in ((0, 500000), (500000, 500003), (500003, T)):
        blk.work(X[:, lo:hi])
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 4000, 0)
        rx.work(X[c])
        for p in range(3):
            got, want = np.concatenate(acc[p][c]), rx.port(p)
            n = min(len(got), len(want))
            assert n > 0 and len(want) - n <= 80, (p, len(got), len(want))
            assert np.array_equal(got[:n], want[:n]), (c, p)
        good, found = siggen.count_good_frames(np.concatenate(acc[2][c]), 0xED89AA, 24, 7, payloads[c])
        assert good >= len(payloads[c]) - 4


def test_carrier_offset_rotator(qrl, oracle):
    """gr_demod_base::set_carrier_offset: per-channel front-end rotator, changed mid-stream with continuous phase."""
    C, T = 3, 300000
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=1500)
    offs = [0.0, 1250.0, -2600.5]
    n = np.arange(T)
    for c in range(C):      # the signal sits `off` Hz above baseband; the receiver rotates it back down (phase inc = -2 pi off / fs)
        X[c] = (X[c] * np.exp(2j * np.pi * offs[c] * n / 1e6)).astype(np.complex64)
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=200000)
    rxs = [oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1) for _ in range(C)]
    for c in range(C):
        blk.set_carrier_offset(offs[c], channel=c)
        rxs[c].set_carrier_offset(offs[c])
    acc = [[[] for _ in range(C)] for _ in range(3)]
    for k, (lo, hi) in enumerate(((0, 150001), (150001, T))):
        if k == 1:          # retune channel 1 mid-stream
            blk.set_carrier_offset(1300.0, channel=1); rxs[1].set_carrier_offset(1300.0)
        blk.work(X[:, lo:hi])
        for c in range(C):
            rxs[c].work(X[c, lo:hi])
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        for p in range(3):
            got, want = np.concatenate(acc[p][c]), rxs[c].port(p)
            nmin = min(len(got), len(want))
            assert nmin > 0 and len(want) - nmin <= 80
            assert np.array_equal(got[:nmin], want[:nmin]), (c, p)


def test_overlapped_calls_give_the_same_stream(qrl, oracle):
    """QRL_PARAM_OVERLAP_CALLS: back-to-back work() calls whose loop / FEC tails overlap the next call's parallel stages
    must produce exactly what the serialised handle (and the oracle) produce, call by call."""
    C, T, K = 3, 1 << 18, 6
    X, _ = siggen.gen_4fsk_channels(C, T * K, seed0=1500)
    ser = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    ovl = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    ovl.set_overlap(True)
    want_calls = []
    for k in range(K):
        ser.work(X[:, k * T:(k + 1) * T])
        want_calls.append([ser.read_port(p) for p in range(3)])
    # (a) read after every call: each read joins the tail
    for k in range(K):
        ovl.work(X[:, k * T:(k + 1) * T])
        got = [ovl.read_port(p) for p in range(3)]
        for p in range(3):
            for c in range(C):
                assert np.array_equal(got[p][c], want_calls[k][p][c]), (k, p, c)
    # (b) a second pass without reading in between: only the last call's ports are read
    ovl2 = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    ovl2.set_overlap(True)
    import torch
    Xd = torch.from_numpy(X).cuda()
    for k in range(K):
        ovl2.work_device(Xd[:, k * T:].data_ptr(), T, Xd.shape[1])
    got = [ovl2.read_port(p) for p in range(3)]
    for p in range(3):
        for c in range(C):
            assert np.array_equal(got[p][c], want_calls[K - 1][p][c]), (p, c)
    # and the concatenation equals the oracle on the whole stream
    bits = [np.concatenate([want_calls[k][2][c] for k in range(K)]) for c in range(C)]
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        rx.work(X[c])
        assert np.array_equal(bits[c], rx.port(2))


def test_overlapped_calls_with_changing_length(qrl, oracle):
    """QRL_PARAM_OVERLAP_CALLS with a different T on every call (the per-slice fences pair slice i of a call with slice i of
    the previous one, which only holds for equal cuts): a call whose length differs joins the previous tail first.  Every call's
    ports must equal the serialised handle's, with and without a read in between, and the bit stream must equal the oracle's."""
    C, Tmax = 3, 1 << 18
    lens = [1 << 18, (1 << 17) + 1234, 1 << 18, 3 << 16, 40000, 1 << 18, 1 << 18, 70001, (1 << 18) - 50]
    X, _ = siggen.gen_4fsk_channels(C, sum(lens), seed0=1700)
    ser = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=Tmax)
    want_calls, o = [], 0
    for T in lens:
        ser.work(np.ascontiguousarray(X[:, o:o + T]))
        want_calls.append([ser.read_port(p) for p in range(3)])
        o += T
    for read_every in (1, 3, len(lens)):
        ovl = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=Tmax)
        ovl.set_overlap(True)
        o = 0
        for k, T in enumerate(lens):
            ovl.work(np.ascontiguousarray(X[:, o:o + T]))
            o += T
            if (k + 1) % read_every == 0 or k == len(lens) - 1:
                got = [ovl.read_port(p) for p in range(3)]
                for p in range(3):
                    for c in range(C):
                        assert np.array_equal(got[p][c], want_calls[k][p][c]), (read_every, k, p, c)
        ovl.close()
    bits = [np.concatenate([want_calls[k][2][c] for k in range(len(lens))]) for c in range(C)]
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        rx.work(X[c])
        assert np.array_equal(bits[c], rx.port(2))


def test_sc8_ingest_equals_the_float_path(qrl):
    """qrl_rx_work_sc8 (int8 I/Q, 2 bytes per sample) against qrl_rx_work on the host-converted stream: identical ports, ragged chunks."""
    C, T = 2, 150000
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=1950)
    q8 = np.empty((C, T, 2), np.int8)
    q8[..., 0] = np.clip(np.round(X.real * 100), -128, 127)
    q8[..., 1] = np.clip(np.round(X.imag * 100), -128, 127)
    scale = np.float32(1.0 / 128.0)
    Xf = (q8[..., 0].astype(np.float32) * scale + 1j * (q8[..., 1].astype(np.float32) * scale)).astype(np.complex64)
    a = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    b = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    lo = 0
    for n in [1, 7, 70001, 4096, 5, 65537, 10353]:
        a.work_sc8(q8[:, lo:lo + n], scale)
        b.work(Xf[:, lo:lo + n])
        for p in range(3):
            for c in range(C):
                assert np.array_equal(a.read_port(p)[c], b.read_port(p)[c]), (lo, p, c)
        lo += n
    assert lo == T


def test_sc16_ingest_equals_the_float_path(qrl, oracle):
    """qrl_rx_work_sc16: int16 I/Q converted on the device (float(v) * scale, one rounding) must give exactly the ports of qrl_rx_work
    fed with the host-converted gr_complex stream, and of the oracle; ragged chunks incl. lengths that break the 16-byte path."""
    C, T = 3, 200000
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=1900)
    q = np.empty((C, T, 2), np.int16)
    q[..., 0] = np.clip(np.round(X.real * 20000), -32768, 32767)
    q[..., 1] = np.clip(np.round(X.imag * 20000), -32768, 32767)
    scale = np.float32(1.0 / 32767.0)
    Xf = (q[..., 0].astype(np.float32) * scale + 1j * (q[..., 1].astype(np.float32) * scale)).astype(np.complex64)
    a = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    b = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    acc_a = [[[] for _ in range(C)] for _ in range(3)]
    acc_b = [[[] for _ in range(C)] for _ in range(3)]
    lo = 0
    for n in [1, 3, 70001, 4096, 2, 65537, 50000, 10360]:
        a.work_sc16(q[:, lo:lo + n], scale)
        b.work(Xf[:, lo:lo + n])
        for p in range(3):
            for c in range(C):
                acc_a[p][c].append(a.read_port(p)[c]); acc_b[p][c].append(b.read_port(p)[c])
        lo += n
    assert lo == T
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_4FSK, 5, 1000000, 1700, 3000, 1)
        rx.work(Xf[c])
        for p in range(3):
            ga, gb = np.concatenate(acc_a[p][c]), np.concatenate(acc_b[p][c])
            assert len(ga) == len(gb) and np.array_equal(ga, gb), (c, p)
            want = rx.port(p)
            assert np.array_equal(ga[:len(want)], want[:len(ga)]) and abs(len(ga) - len(want)) <= 80, (c, p)
