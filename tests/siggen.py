"""Synthetic test signals for the hot path: frames built the way gr_modem::frame()
does (/root/reference/src/gr_modem.cpp:904-961), modulated by the ORACLE's TX chains, then passed through
a simple channel (carrier offset, phase, timing offset, AWGN).  Test infrastructure only."""
import numpy as np

from oracle import oracle as O


def frames_4fsk(rng, n_frames):
    """Voice2 frames: sync 0xED89 + reserved 0xAA + 7 payload bytes (gr_modem.cpp:921-925), 8x0xAA preamble."""
    out = [bytes([0xAA] * 8)]
    payloads = []
    for _ in range(n_frames):
        p = rng.integers(0, 256, 7, dtype=np.uint8).tobytes()
        payloads.append(p)
        out.append(bytes([0xED, 0x89, 0xAA]) + p)
    out.append(bytes([0xAA] * 24))
    return np.frombuffer(b"".join(out), np.uint8).copy(), payloads


def frames_qpsk(rng, n_frames, frame_len=1516):
    """IP frames: sync 0xDE98AA + 1516 payload bytes (gr_modem.cpp:942-947)."""
    out = [bytes([0xAA] * 64)]
    payloads = []
    for _ in range(n_frames):
        p = rng.integers(0, 256, frame_len, dtype=np.uint8).tobytes()
        payloads.append(p)
        out.append(bytes([0xDE, 0x98, 0xAA]) + p)
    out.append(bytes([0xAA] * 128))
    return np.frombuffer(b"".join(out), np.uint8).copy(), payloads


def channel(iq, rng, fo_hz=0.0, phase=0.0, delay=0, snr_db=None, amp=0.8, total=None, fs=1e6):
    x = np.concatenate([np.zeros(delay, np.complex64), iq.astype(np.complex64)])
    if total is not None:
        if len(x) < total:
            x = np.concatenate([x, np.zeros(total - len(x), np.complex64)])
        x = x[:total]
    n = np.arange(len(x))
    x = x * np.exp(1j * (2 * np.pi * fo_hz * n / fs + phase)) * amp
    if snr_db is not None:
        sig_p = amp ** 2 * 0.5
        npow = sig_p / (10 ** (snr_db / 10.0))
        noise = (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x))) * np.sqrt(npow / 2)
        x = x + noise
    return x.astype(np.complex64)


def gen_4fsk_channels(n_channels, T, seed0=1000, snr_db=20.0, fm=True):
    """[C, T] complex64 batch of 4FSK-2k(-FM) bursts + the payloads each channel carries."""
    X = np.zeros((n_channels, T), np.complex64)
    payloads = []
    n_frames = max(1, int((T / 1e6) * 2000 / 8 / 10) - 4)   # 2 kbit/s net, 10-byte frames
    for c in range(n_channels):
        rng = np.random.default_rng(seed0 + c)
        data, pl = frames_4fsk(rng, n_frames)
        tx = O.Tx(O.MOD_4FSK, 25, 1000000, 1700, 3500 if fm else 4000, 1 if fm else 0)
        iq = tx.work(data)
        fo = rng.uniform(-200, 200)
        dl = int(rng.integers(0, 500))
        X[c] = channel(iq, rng, fo_hz=fo, phase=rng.uniform(0, 2 * np.pi), delay=dl, snr_db=snr_db, total=T)
        payloads.append(pl)
    return X, payloads


def gen_qpsk_channels(n_channels, T, seed0=2000, snr_db=15.0):
    X = np.zeros((n_channels, T), np.complex64)
    payloads = []
    n_frames = max(1, int((T / 1e6) * 250000 / 8 / 1519) - 1)
    for c in range(n_channels):
        rng = np.random.default_rng(seed0 + c)
        data, pl = frames_qpsk(rng, n_frames)
        tx = O.Tx(O.MOD_QPSK, 4, 1000000, 1700, 160000, 0)
        iq = tx.work(data)
        fo = rng.uniform(-200, 200)
        X[c] = channel(iq, rng, fo_hz=fo, phase=rng.uniform(0, 2 * np.pi), delay=int(rng.integers(0, 50)),
                       snr_db=snr_db, total=T)
        payloads.append(pl)
    return X, payloads


def count_good_frames(bits, sync, sync_bits, frame_len, payloads):
    fr = O.find_frames(bits, sync, sync_bits, frame_len)
    pl = set(payloads)
    return sum(1 for f in fr if f.tobytes() in pl), len(fr)


def gen_nbfm_channels(n_channels, T, seed0=1, dev_hz=2000.0, snr_sigma=0.01, lead_zeros=0):
    """FM-modulated two-tone audio (1 kHz + 2.2 kHz) at 1 Msps, amplitude 0.8, AWGN sigma."""
    X = np.zeros((n_channels, T), np.complex64)
    n = np.arange(T)
    for c in range(n_channels):
        rng = np.random.default_rng(seed0 + c)
        f1, f2 = 1000.0 + 37.0 * c, 2200.0 - 11.0 * c
        audio = 0.6 * np.sin(2 * np.pi * f1 * n / 1e6) + 0.4 * np.sin(2 * np.pi * f2 * n / 1e6 + 0.3)
        phase = 2 * np.pi * dev_hz * np.cumsum(audio) / 1e6
        x = 0.8 * np.exp(1j * (phase + 2 * np.pi * rng.uniform(-100, 100) * n / 1e6))
        x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * snr_sigma
        if lead_zeros:
            x[:lead_zeros] = 0
        X[c] = x.astype(np.complex64)
    return X
