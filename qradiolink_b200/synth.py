"""Synthetic 1 Msps complex-baseband inputs for benchmarks and demos (no oracle involved).

`burst_4fsk` builds a 4FSK-2k-FM burst the way QRadioLink's transmitter would (frame bytes -> scrambler ->
CCSDS K=7 encoder -> Gray-mapped 4-level symbols -> RRC x25 -> FM -> x20 interpolation), in plain
numpy/scipy double precision.  It is a signal source, not a parity reference: nothing is compared
against it; it only has to look like the traffic the RX chain is built for."""
import ctypes as C

import numpy as np
import scipy.signal as ss

from .lib import load_library


def _taps(fn, *args):
    out = np.zeros(1 << 14, np.float32)
    n = fn(*args, out.ctypes.data_as(C.c_void_p), 1 << 14)
    assert n > 0
    return out[:n].astype(np.float64)


def frame_bytes_4fsk(rng, n_bytes):
    """8x0xAA preamble then Voice2 frames (0xED89 0xAA + 7 payload bytes), padded with 0xAA."""
    out = [0xAA] * 8
    while len(out) + 10 <= n_bytes:
        out += [0xED, 0x89, 0xAA] + list(rng.integers(0, 256, 7))
    out += [0xAA] * (n_bytes - len(out))
    return np.array(out, np.uint8)


def scramble_encode(data_bytes):
    bits = np.unpackbits(data_bytes)
    reg = 0x7F
    coded = np.zeros(2 * len(bits), np.uint8)
    st = 0
    for i, b in enumerate(bits):
        out = reg & 1
        nb = (bin(reg & 0x8A).count("1") & 1) ^ int(b)
        reg = (reg >> 1) | (nb << 7)
        st = ((st << 1) | out) & 0x7F
        coded[2 * i] = bin(st & 109).count("1") & 1
        coded[2 * i + 1] = bin(st & 79).count("1") & 1
    return coded


def burst_4fsk(seed, n_samples):
    L = load_library()
    rng = np.random.default_rng(seed)
    n_bytes = n_samples // 4000 + 1
    coded = scramble_encode(frame_bytes_4fsk(rng, n_bytes))
    chunks = (coded[0::2] << 1) | coded[1::2]
    levels = np.array([-1.5, -0.5, 1.5, 0.5])[chunks]            # map {0,1,3,2} then {-1.5,-0.5,0.5,1.5}
    rrc = _taps(L.qrl_firdes_root_raised_cosine, 25.0, 25.0, 1.0, 0.2, 250)
    shaped = ss.upfirdn(rrc, levels, 25) * 0.66666666
    phase = np.cumsum(shaped * (np.pi / 25))
    iq50 = 0.9 * np.exp(1j * phase)
    lp = _taps(L.qrl_firdes_low_pass, 20.0, 1e6, 3500.0, 3500.0, 0)
    iq = ss.upfirdn(lp, iq50, 20)
    if len(iq) < n_samples:
        iq = np.concatenate([iq, np.zeros(n_samples - len(iq))])
    return iq[:n_samples].astype(np.complex64)


def batch_on_device(bases, n_channels, seed, snr_db=20.0, amp=0.8, device="cuda"):
    """[n_channels, T] complex64 CUDA tensor: channel c = bases[c % len] delayed, frequency/phase shifted, + AWGN."""
    import torch
    T = len(bases[0])
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    base_dev = [torch.from_numpy(b).to(device) for b in bases]
    X = torch.empty((n_channels, T), dtype=torch.complex64, device=device)
    n = torch.arange(T, device=device, dtype=torch.float64)
    sigma = float(np.sqrt(amp * amp * 0.5 / (10 ** (snr_db / 10.0)) / 2))
    for c in range(n_channels):
        fo = rng.uniform(-200, 200)
        ph = rng.uniform(0, 2 * np.pi)
        dl = int(rng.integers(0, 500))
        rot = torch.polar(torch.full((T,), amp, device=device, dtype=torch.float64), 2 * np.pi * fo / 1e6 * n + ph).to(torch.complex64)
        X[c] = torch.roll(base_dev[c % len(base_dev)], dl) * rot
        noise = torch.randn((T, 2), generator=g, device=device, dtype=torch.float32) * sigma
        X[c] += torch.view_as_complex(noise)
    return X
