"""GPU tier: the high-rate FSK modes -- 4FSK10KFM / 2FSK10KFM (rational_resampler_ccf(2,25): 80 ksps) and 4FSK100K
(/2: 500 ksps) -- against the CPU oracle (gr_demod_base.cpp:207,214,225)."""
import numpy as np
import pytest

from tests import siggen

pytestmark = pytest.mark.gpu

CASES = [
    # name, oracle rx kind, rx args, tx kind, tx args, nports
    ("4fsk10k", "DEMOD_4FSK", (1, 1000000, 1700, 20000, 1), "MOD_4FSK", (5, 1000000, 1700, 20000, 1), 3),
    ("2fsk10k", "DEMOD_2FSK", (1, 1000000, 1700, 25000, 1), "MOD_2FSK", (5, 1000000, 1700, 25000, 1), 4),
    ("4fsk100k", "DEMOD_4FSK", (2, 1000000, 1700, 125000, 1), "MOD_4FSK", (2, 1000000, 1700, 125000, 1), 3),
]


@pytest.mark.parametrize("name,rxk,rxa,txk,txa,nports", CASES)
def test_fast_mode_parity(qrl, oracle, name, rxk, rxa, txk, txa, nports):
    C, T = 2, 1 << 19
    rng = np.random.default_rng(hash(name) % 1000)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        nbytes = {"4fsk10k": 1200, "2fsk10k": 1200, "4fsk100k": 12000}[name]
        data = rng.integers(0, 256, nbytes, dtype=np.uint8)
        iq = oracle.Tx(getattr(oracle, txk), *txa).work(data)
        X[c] = siggen.channel(iq, rng, fo_hz=rng.uniform(-100, 100), delay=int(rng.integers(0, 100)), snr_db=20.0, amp=0.1, total=T)
    make = qrl.make_gr_demod_4fsk if rxk == "DEMOD_4FSK" else qrl.make_gr_demod_2fsk
    blk = make(rxa[0], rxa[1], rxa[2], rxa[3], bool(rxa[4]), n_channels=C, max_samples=300000)
    acc = [[[] for _ in range(C)] for _ in range(nports)]
    for lo, hi in ((0, 300000), (300000, 300013), (300013, T)):
        blk.work(X[:, lo:hi])
        for p in range(nports):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(C):
        rx = oracle.Rx(getattr(oracle, rxk), *rxa)
        rx.work(X[c])
        for p in range(nports):
            got, want = np.concatenate(acc[p][c]), rx.port(p)
            n = min(len(got), len(want))
            assert n > 100 and len(want) - n <= 160, (p, len(got), len(want))
            assert np.array_equal(got[:n], want[:n]), (name, c, p)
