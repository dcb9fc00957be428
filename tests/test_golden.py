"""Known-answer tests against the committed vectors in tests/golden/golden_v1.json (made by
tests/golden/make_golden.py).  CPU tier: the oracle must still reproduce them (and, where the reference tree is
available, so must the reference's own emphasis.cpp).  GPU tier: the CUDA path, through the C ABI, must reproduce
them bit for bit -- decoded bits, counts and the SHA-256 of every float port."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import siggen
from tests.golden import cases

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_stream(got, want, what):
    got = np.asarray(got)
    assert len(got) == want["n"], (what, len(got), want["n"])
    if "hex" in want:
        packed = np.packbits(got).tobytes().hex() if len(got) and got.max() <= 1 else got.tobytes().hex()
        assert packed == want["hex"], what
    else:
        v = got.view(np.float32) if got.dtype == np.complex64 else got.astype(np.float32)
        assert [float(x) for x in v[:8]] == want["head"], what
        assert sha(got) == want["sha256"], what


# ------------------------------------------------------------------------------------------------ CPU tier
def test_golden_emphasis_reference_and_oracle(oracle):
    """The committed taps came from the reference's emphasis.cpp; the oracle's restatement must equal them (bitwise),
    and so must the reference itself wherever oracle/_ref exists."""
    R = cases.load_ref_emphasis()
    for fs, tau in cases.EMPHASIS_CASES:
        wd = [float.fromhex(x) for x in G["emphasis"]["deemph_%d_%g" % (fs, tau)]]
        wp = [float.fromhex(x) for x in G["emphasis"]["preemph_%d_%g" % (fs, tau)]]
        a, b = oracle.deemph_taps(fs, tau)
        assert list(a) + list(b) == wd
        a, b = oracle.preemph_taps(fs, tau)
        assert list(a) + list(b) == wp
        if R is not None:
            assert cases.ref_deemph(R, fs, tau) == wd and cases.ref_preemph(R, fs, tau) == wp


def test_golden_design_and_kat_oracle(oracle):
    for name, fn in cases.DESIGN_CASES.items():
        t = np.asarray(fn(oracle), np.float32)
        assert len(t) == G["design"][name]["n"] and sha(t) == G["design"][name]["sha256"], name
    k = G["kat"]
    bits = np.unpackbits(np.frombuffer(bytes.fromhex(k["bits_hex"]), np.uint8))[:400]
    enc = oracle.cc_encode(bits)
    assert np.packbits(enc).tobytes().hex() == k["cc_encode_hex"]
    assert np.packbits(oracle.cc_decode((enc.astype(np.int32) * 255).astype(np.uint8))).tobytes().hex() == k["cc_decode_of_encoded_hex"]
    assert np.packbits(oracle.scramble(bits)).tobytes().hex() == k["scramble_hex"]
    assert np.packbits(oracle.descramble(bits)).tobytes().hex() == k["descramble_hex"]


@pytest.mark.parametrize("name", ["4fsk_2k_fm", "nbfm_2500"])
def test_golden_rx_oracle(oracle, name):
    """Two of the RX cases on the CPU (the rest run in the GPU tier, where the oracle and CUDA are both checked)."""
    case = cases.RX_CASES[name]
    X = case["signal"](oracle, siggen)
    assert sha(X) == G["rx"][name]["input_sha256"], "signal generator drifted"
    for c in range(X.shape[0]):
        rx = oracle.Rx(case["okind"], *case["args"])
        rx.work(X[c])
        for p in range(case["nports"]):
            check_stream(rx.port(p), G["rx"][name]["channels"][c][p], (name, c, p))


def test_golden_tx_oracle(oracle):
    for name, case in cases.TX_CASES.items():
        data = case["data"]()
        assert sha(data) == G["tx"][name]["input_sha256"]
        check_stream(oracle.Tx(case["okind"], *case["args"]).work(data), G["tx"][name]["out"], name)


def test_golden_extra_blocks_oracle(oracle):
    """MMDVM channel chains, display spectrum, DSSS despreader: the oracle against the committed answers."""
    for name, a in cases.extra_outputs(oracle).items():
        check_stream(a, G["extra"][name], name)


# ------------------------------------------------------------------------------------------------ GPU tier
@pytest.mark.gpu
def test_golden_mmdvm_chains_cuda(qrl):
    """The MMDVM channel chains through the C ABI against the committed answers (int16 and IQ are bit-exact on the device)."""
    rx = qrl.MmdvmChannelsRx(1, filter_width=5000, max_in=30000)
    o, db, at = rx.work(cases.mmdvm_rx_input()[None, :])
    check_stream(o[0], G["extra"]["mmdvm_rx_int16"], "mmdvm_rx_int16")
    assert len(db[0]) == G["extra"]["mmdvm_rx_rssi_db"]["n"]
    assert np.max(np.abs(db[0][:8] - np.array(G["extra"]["mmdvm_rx_rssi_db"]["head"], np.float32)[:len(db[0][:8])])) < 2e-4
    tx = qrl.MmdvmChannelsTx(1, filter_width=5000, max_in=12000)
    check_stream(tx.work(cases.mmdvm_tx_input()[None, :])[0], G["extra"]["mmdvm_tx_iq"], "mmdvm_tx_iq")



@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.RX_CASES))
def test_golden_rx_cuda(qrl, oracle, name):
    case = cases.RX_CASES[name]
    X = case["signal"](oracle, siggen)
    assert sha(X) == G["rx"][name]["input_sha256"], "signal generator drifted"
    blk = getattr(qrl, case["factory"])(*case["fargs"], n_channels=X.shape[0], max_samples=X.shape[1])
    # two calls with an odd split: the answer may not depend on chunking
    cut = min(100003, X.shape[1] // 2 + 3)
    acc = [[[] for _ in range(X.shape[0])] for _ in range(case["nports"])]
    for lo, hi in ((0, cut), (cut, X.shape[1])):
        blk.work(X[:, lo:hi])
        for p in range(case["nports"]):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    for c in range(X.shape[0]):
        for p in range(case["nports"]):
            check_stream(np.concatenate(acc[p][c]), G["rx"][name]["channels"][c][p], (name, c, p))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.TX_CASES))
def test_golden_tx_cuda(qrl, name):
    case = cases.TX_CASES[name]
    data = case["data"]()
    tx = getattr(qrl, case["factory"])(*case["fargs"], n_channels=1, max_items=len(data))
    check_stream(tx.work(data[None, :])[0], G["tx"][name]["out"], name)


def test_golden_design_product_host_side(qrl):
    """qrl_design_deemph (the product's host-side design code, no GPU needed) against the reference-derived taps."""
    import ctypes as C
    L = qrl.load_library()
    for fs, tau in cases.EMPHASIS_CASES:
        a = np.zeros(2); b = np.zeros(2)
        assert L.qrl_design_deemph(fs, tau, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == 0
        assert list(a) + list(b) == [float.fromhex(x) for x in G["emphasis"]["deemph_%d_%g" % (fs, tau)]]
