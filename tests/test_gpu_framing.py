"""GPU tier: the device deframer (gr_modem::synchronize / findSync / packBytes) against the oracle's
bit-serial restatement: identical frame lists for all three sync classes, random bit streams with planted and accidental
sync words, frames straddling chunk boundaries, record overflow, and end to end behind the CUDA demodulator."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits_of(bs):
    return np.unpackbits(np.frombuffer(bytes(bs), np.uint8))


def planted_stream(rng, sync_class, bit_buf_len, n_frames):
    words = {1: [[0xB5]], 2: [[0xED, 0x89], [0x89, 0xED, 0xAA], [0xED, 0x77, 0xAA], [0x98, 0xDE, 0xAA], [0x8C, 0xC8, 0xDD], [0x4C, 0x8A, 0x2B]],
             3: [[0xDE, 0x98, 0xAA], [0x98, 0xDE, 0xAA], [0x4C, 0x8A, 0x2B]]}[sync_class]
    parts = [rng.integers(0, 2, int(rng.integers(0, 40)), dtype=np.uint8)]
    for _ in range(n_frames):
        w = words[int(rng.integers(0, len(words)))]
        parts.append(bits_of(w))
        parts.append(rng.integers(0, 2, bit_buf_len, dtype=np.uint8))        # enough payload bits for any type
        parts.append(rng.integers(0, 2, int(rng.integers(0, 70)), dtype=np.uint8))
    return np.concatenate(parts)


@pytest.mark.parametrize("sync_class,bit_buf_len,rx_len", [(1, 32, 4), (2, 64, 7), (2, 384, 47), (3, 623 * 8, 622)])
def test_deframer_matches_oracle(qrl, oracle, sync_class, bit_buf_len, rx_len):
    C = 5
    rng = np.random.default_rng(40 + sync_class + bit_buf_len)
    streams = [planted_stream(rng, sync_class, bit_buf_len, 6 + c) for c in range(C)]
    streams[1] = rng.integers(0, 2, 20000, dtype=np.uint8)                    # pure noise: accidental sync words
    streams[2] = np.zeros(0, np.uint8)                                        # empty channel
    want = []
    for s in streams:
        want.append(oracle.Deframer(sync_class, bit_buf_len, rx_len).work(s))
    d = qrl.Deframer(sync_class, bit_buf_len, rx_len, n_channels=C, max_bits=max(len(s) for s in streams) + 8)
    got = d.work(streams)
    assert got == want
    assert sum(len(f) for f in got) >= 10
    # the same streams in ragged chunks: partial frames and shift registers carry over
    d2 = qrl.Deframer(sync_class, bit_buf_len, rx_len, n_channels=C, max_bits=4096)
    o2 = [oracle.Deframer(sync_class, bit_buf_len, rx_len) for _ in range(C)]
    acc = [[] for _ in range(C)]; acc_o = [[] for _ in range(C)]
    lo, i, sizes = 0, 0, [1, 31, 32, 33, 4096, 7, 1000, 64]
    n = max(len(s) for s in streams)
    while lo < n:
        step = sizes[i % len(sizes)]; i += 1
        chunk = [s[lo:lo + step] for s in streams]
        for c, fr in enumerate(d2.work(chunk)):
            acc[c] += fr
        for c in range(C):
            acc_o[c] += o2[c].work(chunk[c])
        lo += step
    assert acc == want and acc_o == want
    assert list(d2.modem_sync) == [o.modem_sync for o in o2]


def test_deframer_record_overflow_drops_like_the_oracle(qrl, oracle):
    rng = np.random.default_rng(3)
    s = planted_stream(rng, 2, 64, 30)
    d = qrl.Deframer(2, 64, 7, n_channels=1, max_bits=len(s), max_frames=4)
    got = d.work([s])[0]
    assert len(got) == 4 and got == oracle.Deframer(2, 64, 7).work(s)[:4]


def test_deframer_behind_the_demodulator_on_device(qrl, oracle):
    """4FSK-2k-FM RX -> port 2 stays on the GPU -> deframer: the voice frames' payloads are the transmitted ones."""
    from tests import siggen
    C, T = 3, 1 << 19
    X, payloads = siggen.gen_4fsk_channels(C, T, seed0=8800)
    rx = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    rx.work(X)
    d = qrl.Deframer.for_mode("4FSK2KFM", n_channels=C, max_bits=1 << 16)
    frames = d.work_from_rx(rx, port=2)
    bits = rx.read_port(2)
    for c in range(C):
        assert frames[c] == oracle.Deframer(2, 64, 7).work(bits[c])
        voice = [p[1:] for ty, p in frames[c] if ty == 0xED89 and len(p) == 8]
        assert len(set(voice) & set(payloads[c])) >= 3


@pytest.mark.parametrize("modem_type", [1, 2, 3])
def test_gr_deframer_bb_on_device_matches_oracle(qrl, oracle, modem_type):
    """gr_deframer_bb (gr_deframer_bb.cpp:83-185) for a batch of channels against the oracle restatement (itself pinned to the
    reference source, tests/test_oracle_ref.py): one shot and in ragged chunks, sync words straddling chunk boundaries."""
    C = 4
    rng = np.random.default_rng(50 + modem_type)
    words = [(0xED89, 16), (0x89ED, 16), (0x98DE, 16), (0xED77, 16), (0x8CC8, 16), (0x4C8A2B, 24), (0xB5, 8)]
    streams = []
    for c in range(C):
        n = 30000 + 1111 * c
        bits = rng.integers(0, 2, n, dtype=np.uint8)
        pos = 17
        while pos + 500 < n:
            w, nb = words[int(rng.integers(0, len(words)))]
            bits[pos:pos + nb] = [(w >> (nb - 1 - k)) & 1 for k in range(nb)]
            pos += int(rng.integers(60, 700))
        streams.append(bits)
    streams[2] = np.zeros(0, np.uint8)
    want = [oracle.DeframerBB(modem_type).work(s) for s in streams]
    d = qrl.DeframerBB(modem_type, n_channels=C, max_bits=max(len(s) for s in streams) + 8)
    got = d.work(streams)
    for c in range(C):
        assert np.array_equal(got[c], want[c]), c
    assert sum(len(g) for g in got) > 2000
    d2 = qrl.DeframerBB(modem_type, n_channels=C, max_bits=4096)
    acc = [[] for _ in range(C)]
    lo, i, sizes = 0, 0, [1, 31, 32, 33, 4096, 7, 1000, 64, 15]
    n = max(len(s) for s in streams)
    while lo < n:
        step = sizes[i % len(sizes)]; i += 1
        for c, o in enumerate(d2.work([s[lo:lo + step] for s in streams])):
            acc[c].append(o)
        lo += step
    for c in range(C):
        assert np.array_equal(np.concatenate(acc[c]) if acc[c] else np.zeros(0, np.uint8), want[c]), c


def test_dual_decoder_chain_stays_on_the_device(qrl, oracle):
    """BPSK-2k RX -> ports 2 / 3 -> two gr_deframer_bb -> the longer stream -> gr_modem::synchronize, all on the GPU
    (gr_demod_base.cpp _deframer1/2 + gr_modem.cpp:1043-1095), against the same chain of oracle restatements."""
    from tests import siggen
    C, T = 2, 1 << 19
    rng = np.random.default_rng(9100)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        data, _ = siggen.frames_4fsk(rng, 18)           # 0xED89 voice frames: the BPSK-2k framing is the same narrow class
        iq = oracle.Tx(oracle.MOD_BPSK, 250, 1000000, 1700, 2400, 0).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-100, 100), phase=rng.uniform(0, 6.28), delay=int(rng.integers(0, 300)), snr_db=22.0, total=T)
    rx = qrl.make_gr_demod_bpsk(5, 1000000, 1700, 2400, n_channels=C, max_samples=T)
    rx.work(X)
    b2, b3 = rx.read_port(2), rx.read_port(3)
    cap = max(max(len(b) for b in b2), max(len(b) for b in b3)) + 64
    da, db = qrl.DeframerBB(1, n_channels=C, max_bits=1 << 16), qrl.DeframerBB(1, n_channels=C, max_bits=1 << 16)
    oa, ob = da.work_from_rx(rx, 2), db.work_from_rx(rx, 3)
    fr = qrl.Deframer.for_mode("BPSK2K", n_channels=C, max_bits=(1 << 16) + 32)
    frames = fr.work2_from_dfbb(da, db)
    total = 0
    for c in range(C):
        wa, wb = oracle.DeframerBB(1).work(b2[c]), oracle.DeframerBB(1).work(b3[c])
        assert np.array_equal(oa[c], wa) and np.array_equal(ob[c], wb)
        pick = wa if len(wa) >= len(wb) else wb
        assert frames[c] == oracle.Deframer(2, 64, 7).work(pick)
        total += len(frames[c])
    assert total >= 10 and cap > 0
    # host-buffer form of the same selection
    assert qrl.Deframer.for_mode("BPSK2K", n_channels=C, max_bits=(1 << 16) + 32).work2(oa, ob) == frames


def test_deframer_reports_dropped_frames_and_reverts_streams(qrl):
    rng = np.random.default_rng(3)
    s = planted_stream(rng, 2, 64, 30)
    d = qrl.Deframer(2, 64, 7, n_channels=1, max_bits=len(s), max_frames=4)
    assert len(d.work([s])[0]) == 4 and int(d.dropped()[0]) >= 20
    d.set_stream(0)                                   # NULL: back to a stream of its own (include/qrl_b200.h)
    assert len(d.work([s[:10]])[0]) == 0


def test_m17_sync_class_matches_oracle(qrl, oracle):
    """ModemTypeM17 branch of findSync (gr_modem.cpp:1187-1207): link-setup / stream words and the 32-bit end-of-transmission word."""
    rng = np.random.default_rng(77)
    C = 3
    streams = []
    for c in range(C):
        parts = [rng.integers(0, 2, 13 + c, dtype=np.uint8)]
        for k in range(7):
            w = [[0x55, 0xF7], [0xFF, 0x5D], [0x55, 0x5D, 0x55, 0x5D]][int(rng.integers(0, 3))]
            parts += [bits_of(w), rng.integers(0, 2, 46 * 8, dtype=np.uint8), rng.integers(0, 2, int(rng.integers(0, 50)), dtype=np.uint8)]
        streams.append(np.concatenate(parts))
    want = [oracle.Deframer(4, 46 * 8, 46).work(s) for s in streams]
    assert sum(len(w) for w in want) >= 15 and {ty for w in want for ty, _ in w} >= {0x55F7, 0xFF5D}
    d = qrl.Deframer.for_mode("M17", n_channels=C, max_bits=max(len(s) for s in streams))
    assert d.work(streams) == want
    d2 = qrl.Deframer.for_mode("M17", n_channels=C, max_bits=512)
    acc = [[] for _ in range(C)]
    lo, n = 0, max(len(s) for s in streams)
    while lo < n:
        for c, fr in enumerate(d2.work([s[lo:lo + 97] for s in streams])):
            acc[c] += fr
        lo += 97
    assert acc == want


def test_tx_frame_matches_oracle_and_round_trips(qrl, oracle):
    """gr_modem::frame (gr_modem.cpp:904-961) on the device against its oracle restatement, then frame -> CUDA 4FSK TX -> CUDA RX ->
    device deframer recovers the payloads."""
    rng = np.random.default_rng(78)
    types = [0xED89, 0x89EDAA, 0xDE98AA, 0x98DEAA, 0xED77AA, 0x8CC8DD, 0x4C8A2B]
    payloads = [rng.integers(0, 256, int(rng.integers(1, 60)), dtype=np.uint8).tobytes() for _ in types]
    for one_k in (False, True):
        for burst in (False, True):
            got = qrl.frame(payloads, types, one_k_mode=one_k, burst_ip=burst)
            for g, p, t in zip(got, payloads, types):
                assert np.array_equal(g, oracle.frame(p, t, one_k, burst)), (hex(t), one_k, burst)
    C = 2
    pl = [[rng.integers(0, 256, 7, dtype=np.uint8).tobytes() for _ in range(20)] for _ in range(C)]
    data = []
    for c in range(C):
        frames = qrl.frame(pl[c], [0xED89] * 20)
        data.append(np.concatenate([np.full(8, 0xAA, np.uint8)] + frames + [np.full(24, 0xAA, np.uint8)]))
    tx = qrl.make_gr_mod_4fsk(25, 1000000, 1700, 3500, True, n_channels=C, max_items=len(data[0]))
    iq = tx.work(np.stack(data)) * 0.8
    rx = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=iq.shape[1])
    rx.work(iq)
    fr = qrl.Deframer.for_mode("4FSK2KFM", n_channels=C, max_bits=1 << 16).work_from_rx(rx, port=2)
    for c in range(C):
        voice = [p[1:] for ty, p in fr[c] if ty == 0xED89 and len(p) == 8]
        assert len(set(voice) & set(pl[c])) >= 15
