"""Generates tests/golden/golden_v1.json -- known-answer vectors for the hot path.

Where they come from (the reference ships no golden vectors and its arithmetic lives in
un-vendored GNU Radio / VOLK, so parity against real GNU Radio stays UNPINNED):

  * "emphasis": de-/pre-emphasis IIR taps computed by the REFERENCE's own src/gr/emphasis.cpp, compiled from
    /root/reference by oracle/Makefile into oracle/_ref/libqrl_ref_emphasis.so.  This is the one piece of the path
    that is pinned to reference code; the values are committed so the GPU box (which has no /root/reference) can
    check them too.
  * everything else: outputs of the CPU oracle (oracle/qrl_oracle.c) on seeded inputs, frozen here so that
    (a) any later change of the oracle's arithmetic is caught by `-m "not gpu"` tests, and
    (b) the CUDA path is checked against committed vectors, not only against whatever the oracle computes today.
    Float streams are recorded as SHA-256 of their bytes + a few leading values; bits as packed hex.

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import siggen  # noqa: E402
from tests.golden import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe(a):
    a = np.asarray(a)
    if a.dtype == np.uint8:
        return {"n": int(len(a)), "hex": np.packbits(a).tobytes().hex() if len(a) and a.max() <= 1 else a.tobytes().hex()}
    v = a.view(np.float32) if a.dtype == np.complex64 else a.astype(np.float32)          # (the hash is over the raw bytes, int16 included)
    return {"n": int(len(a)), "sha256": sha(a), "head": [float(x) for x in v[:8]]}


def main():
    out = {"version": 1, "generator": "tests/golden/make_golden.py", "rx": {}, "tx": {}, "kat": {}, "design": {}}

    # ---- reference-pinned: emphasis.cpp
    ref = cases.load_ref_emphasis()
    if ref is None:
        raise SystemExit("oracle/_ref/libqrl_ref_emphasis.so missing: run `make -C oracle` where /root/reference exists")
    out["emphasis"] = {}
    for fs, tau in cases.EMPHASIS_CASES:
        out["emphasis"]["deemph_%d_%g" % (fs, tau)] = [float.hex(x) for x in cases.ref_deemph(ref, fs, tau)]
        out["emphasis"]["preemph_%d_%g" % (fs, tau)] = [float.hex(x) for x in cases.ref_preemph(ref, fs, tau)]

    # ---- oracle: design functions
    for name, fn in cases.DESIGN_CASES.items():
        t = fn(O)
        out["design"][name] = {"n": int(len(t)), "sha256": sha(np.asarray(t, np.float32))}

    # ---- oracle: integer known-answer tests
    rng = np.random.default_rng(7)
    bits = rng.integers(0, 2, 400, dtype=np.uint8)
    enc = O.cc_encode(bits)
    soft = (enc.astype(np.int32) * 255).astype(np.uint8)
    out["kat"]["bits_hex"] = np.packbits(bits).tobytes().hex()
    out["kat"]["cc_encode_hex"] = np.packbits(enc).tobytes().hex()
    out["kat"]["cc_decode_of_encoded_hex"] = np.packbits(O.cc_decode(soft)).tobytes().hex()
    out["kat"]["scramble_hex"] = np.packbits(O.scramble(bits)).tobytes().hex()
    out["kat"]["descramble_hex"] = np.packbits(O.descramble(bits)).tobytes().hex()

    # ---- oracle: RX chains on seeded signals
    for name, case in cases.RX_CASES.items():
        X = case["signal"](O, siggen)
        rec = {"input_sha256": sha(X), "channels": []}
        for c in range(X.shape[0]):
            rx = O.Rx(case["okind"], *case["args"])
            rx.work(X[c])
            rec["channels"].append([describe(rx.port(p)) for p in range(case["nports"])])
        out["rx"][name] = rec

    # ---- oracle: TX chains
    for name, case in cases.TX_CASES.items():
        data = case["data"]()
        tx = O.Tx(case["okind"], *case["args"])
        y = tx.work(data)
        out["tx"][name] = {"input_sha256": sha(data), "out": describe(y)}

    # ---- oracle: blocks outside the Rx / Tx factories (MMDVM channel chains, display spectrum, DSSS despreader)
    out["extra"] = {name: describe(a) for name, a in cases.extra_outputs(O).items()}

    path = os.path.join(ROOT, "tests", "golden", "golden_v1.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
