"""CPU tier: the C-ABI library loads and exports every symbol include/qrl_b200.h declares; host-only design
helpers agree bit-for-bit with the oracle's; creating a handle without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(qrl):
    hdr = open(os.path.join(ROOT, "include", "qrl_b200.h")).read()
    declared = set(re.findall(r"\b(qrl_[a-z0-9_]+)\s*\(", hdr))
    from qradiolink_b200 import lib
    L = lib.load_library()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(lib.SYMBOLS), declared ^ set(lib.SYMBOLS)


def test_no_cpu_fallback(qrl):
    if qrl.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(qrl.QrlError) as e:
        qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True)
    assert "no CUDA device" in str(e.value)


def _taps(fn, *args, per=1):
    out = np.zeros(1 << 15, np.float32)
    n = fn(*args, out.ctypes.data_as(C.c_void_p), (1 << 15) // per)
    assert n > 0
    return out[: n * per]


def test_design_matches_oracle_bitwise(qrl, oracle):
    L = qrl.load_library()
    O = oracle
    BH = O.WIN_BLACKMAN_HARRIS
    cases = [
        (_taps(L.qrl_firdes_low_pass, 1.0, 1e6, 10000.0, 10000.0, BH), O.low_pass(1, 1e6, 10000, 10000, BH)),
        (_taps(L.qrl_firdes_low_pass, 1.0, 20000.0, 3000.0, 1500.0, BH), O.low_pass(1, 20000, 3000, 1500, BH)),
        (_taps(L.qrl_firdes_low_pass, 20.0, 1e6, 3500.0, 3500.0, O.WIN_HAMMING), O.low_pass(20, 1e6, 3500, 3500, O.WIN_HAMMING)),
        (_taps(L.qrl_firdes_low_pass_2, 1.0, 1e6, 250000.0, 50000.0, 60.0, BH), O.low_pass_2(1, 1e6, 250000, 50000, 60, BH)),
        (_taps(L.qrl_firdes_root_raised_cosine, 1.5, 20000.0, 2000.0, 0.2, 251), O.rrc(1.5, 20000, 2000, 0.2, 251)),
        (_taps(L.qrl_firdes_root_raised_cosine, 2.0, 2.0, 1.0, 0.35, 22), O.rrc(2, 2, 1, 0.35, 22)),
        (_taps(L.qrl_firdes_gaussian, 50.0, 50.0, 0.3, 55), O.gaussian(50, 50, 0.3, 55)),
        (_taps(L.qrl_firdes_gaussian, 100.0, 100.0, 0.3, 35), O.gaussian(100, 100, 0.3, 35)),
        (_taps(L.qrl_firdes_band_pass, 1.0, 8000.0, 300.0, 3500.0, 200.0, BH), O.band_pass(1, 8000, 300, 3500, 200, BH)),
        (_taps(L.qrl_firdes_complex_band_pass, 1.0, 20000.0, -4000.0, -2000.0, 4000.0, BH, per=2),
         O.complex_band_pass(1, 20000, -4000, -2000, 4000, BH).view(np.float32)),
    ]
    for got, want in cases:
        assert len(got) == len(want)
        assert np.array_equal(got.view(np.uint32), np.asarray(want, np.float32).view(np.uint32))
    for name in ("atan", "tanh", "mmse", "fxpt_sine"):
        want = O.table(name)
        got = np.zeros(len(want), np.float32)
        assert L.qrl_design_table(name.encode(), got.ctypes.data_as(C.c_void_p), len(got)) == len(want)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    a = np.zeros(2); b = np.zeros(2)
    L.qrl_design_deemph(20000, 50e-6, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    oa, ob = O.deemph_taps(20000, 50e-6)
    assert np.array_equal(a, oa) and np.array_equal(b, ob)


def test_sink_semantics(qrl):
    s = qrl.gr_bit_sink()
    s.work(np.ones(31, np.uint8))
    assert s.get_data() is None                       # < 32 items -> nullptr (gr_bit_sink.cpp:49-52)
    s.work(np.ones(1, np.uint8))
    assert len(s.get_data()) == 32 and s.get_data() is None
    a = qrl.gr_audio_sink()
    a.work(np.zeros(639, np.float32))
    assert a.get_data() is None
    a.work(np.zeros(700, np.float32))
    assert len(a.get_data()) == 640 and len(a.get_data()) == 640 and a.get_data() is None
    a.work(np.zeros(8001, np.float32))
    a.work(np.zeros(10, np.float32))                  # > 8000 pending -> buffer cleared (gr_audio_sink.cpp:77-83)
    assert a.get_data() is None
