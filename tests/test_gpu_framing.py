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
