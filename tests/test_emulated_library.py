"""CPU tier: the library's OWN sources (qradiolink_b200/csrc/*.cu, *.cuh) compiled for host threads by tools/emu/build_emulated_lib.py
and driven through the real C ABI and the real Python wrappers -- handle creation, buffer sizing, launch geometry, slice pipelining,
the warp-specialised TMA producer / consumer kernels, the lean symbol-sync recurrence -- against the oracle and the committed golden
vectors.  The test bodies are the GPU tier's own (tests/test_golden.py, tests/test_gpu_*.py), called with the emulated build.

TEST INFRASTRUCTURE: the emulated library is built into a temporary directory, patched into the ctypes
loader for this module only and removed again; the package has no switch for it and no CPU fallback.  It says nothing about timing,
races on real hardware or the memory model -- that is the GPU tier's job; what it does catch is wrong indexing, wrong buffer sizes,
wrong launch geometry, state that does not carry across calls, and barrier mistakes (it found a missing __syncthreads in the SSB
audio kernel that the GPU's scheduling had been hiding)."""
import os
import sys

import numpy as np
import pytest

from tests.golden import cases

pytestmark = pytest.mark.timeout(900)      # pytest-timeout: a scheduling bug in the emulator must fail, not hang the tier

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu_qrl(tmp_path_factory):
    sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))
    from build_emulated_lib import build
    lib_path, n_sites = build(str(tmp_path_factory.mktemp("qrl_emu")))
    assert n_sites >= 60                                   # every kernel<<<...>>> launch site of the three translation units
    import qradiolink_b200 as q
    from qradiolink_b200 import lib as L
    saved = (L._LIB_PATH, L._LIB)
    L._LIB_PATH, L._LIB = lib_path, None
    q.load_library()
    assert L.device_count() == 1
    yield q
    L._LIB_PATH, L._LIB = saved


# the cheaper golden cases (the 250 ksps QPSK chain takes minutes under emulation: GPU tier only)
@pytest.mark.parametrize("name", ["4fsk_2k_fm", "nbfm_2500", "am_5000", "bpsk_2k", "2fsk_2k_fm", "ssb_usb", "wbfm_75k", "m17"])
def test_golden_rx_through_the_emulated_library(emu_qrl, oracle, name):
    from tests import test_golden as TG
    assert name in cases.RX_CASES
    TG.test_golden_rx_cuda(emu_qrl, oracle, name)


@pytest.mark.parametrize("name", sorted(cases.TX_CASES))
def test_golden_tx_through_the_emulated_library(emu_qrl, name):
    from tests import test_golden as TG
    TG.test_golden_tx_cuda(emu_qrl, name)


def test_dmr_chain_through_the_emulated_library(emu_qrl, oracle):
    """QRL_DEMOD_DMR has not run on a GPU yet (tests/test_gpu_dmr.py is opt-in): here its real host path -- create, x3/125 generic
    stage 1, 1-tap stage 2, float port 3, symbol-sync variant 3, scaling epilogue, hard bits -- runs against the oracle."""
    from tests import test_gpu_dmr as TD
    TD.test_dmr_parity_chunked(emu_qrl, oracle)
    TD.test_dmr_refuses_overlapped_calls(emu_qrl)


def test_m17_modulator_through_the_emulated_library(emu_qrl, oracle):
    """QRL_MOD_M17 (tests/test_gpu_m17_tx.py is opt-in until a GPU run): real host path, state carried across three calls."""
    C, n = 2, 48
    data = np.random.default_rng(5400).integers(0, 256, (C, n), dtype=np.uint8)
    tx = emu_qrl.make_gr_mod_m17(n_channels=C, max_items=n)
    got = np.concatenate([tx.work(data[:, a:b]) for a, b in ((0, 1), (1, 20), (20, n))], axis=1)
    for c in range(C):
        want = oracle.Tx(oracle.MOD_M17, 125, 1000000, 1700, 9000, 0).work(data[c])
        assert got.shape[1] == len(want) == n * 4 * 5 * 125 // 3
        assert np.array_equal(got[c], want), c


def test_overlapped_calls_through_the_emulated_library(emu_qrl, oracle):
    """QRL_PARAM_OVERLAP_CALLS (double-buffered ports, ring fences, tail joins): same stream as serialised calls and as the oracle."""
    from tests import siggen
    from tests.test_gpu_4fsk import run_oracle, compare
    C, T = 3, 150000
    X, _ = siggen.gen_4fsk_channels(C, T, seed0=1000)
    want = run_oracle(oracle, X)
    blk = emu_qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=70001)
    blk.set_overlap(True)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    lo, i, sizes = 0, 0, [1, 49, 50, 51, 70001, 12345, 33333, 7]
    while lo < T:
        n = min(sizes[i % len(sizes)], T - lo); i += 1
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    compare([[np.concatenate(acc[p][c]) for c in range(C)] for p in range(3)], want, exact_float=True)


def test_channelizer_and_deframer_through_the_emulated_library(emu_qrl, oracle):
    from tests import test_gpu_pfb as TP
    from tests import test_gpu_framing as TF
    TP.test_channelizer_reference_config_bit_identical_and_chunked(emu_qrl, oracle)
    TP.test_synthesizer_matches_oracle_and_loops_back(emu_qrl, oracle)
    TF.test_deframer_record_overflow_drops_like_the_oracle(emu_qrl, oracle)


@pytest.mark.parametrize("single", [False, True])
def test_mmdvm_tx_zero_idle_through_the_emulated_library(emu_qrl, oracle, single):
    """gr_zero_idle_bursts(0) on the MMDVM modulators (qrl_mmdvm_tx_zero_samples): host bookkeeping of the tags + the clearing kernel
    between the stages, against the oracle, before GPU time is spent on it."""
    from tests import test_gpu_mmdvm as TM
    TM.zero_idle_case(emu_qrl, oracle, single, q=1)


def test_qpsk_chain_through_the_emulated_library(emu_qrl, oracle):
    """The QPSK-250k chain's loop kernels hand blocks between warps through 3-deep shared-memory rings (AGC -> Costas -> bulk store;
    symbol sync -> second Costas -> feed-forward epilogue): a short ragged stream through the real host code and kernels, ports
    bit-identical to the oracle.  (The golden QPSK case and the GPU tier's QPSK tests also pass under emulation; they take minutes.)"""
    from tests import siggen
    C, T = 2, 36000
    X, _ = siggen.gen_qpsk_channels(C, T, seed0=2300)
    blk = emu_qrl.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=20001)
    acc = [[[] for _ in range(C)] for _ in range(3)]
    lo = 0
    for n in (1, 20001, 3, T - 20005):
        blk.work(X[:, lo:lo + n]); lo += n
        for p in range(3):
            for c, v in enumerate(blk.read_port(p)):
                acc[p][c].append(v)
    assert lo == T
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_QPSK, 2, 1000000, 1700, 160000, 0)
        rx.work(X[c])
        for p in range(3):
            g, w = np.concatenate(acc[p][c]), rx.port(p)
            n = min(len(g), len(w))
            assert n >= len(w) - (160 if p == 2 else 8), (c, p, len(g), len(w))
            assert np.array_equal(g[:n], w[:n]), (c, p)
        assert len(np.concatenate(acc[2][c])) > 3000           # several Viterbi frames came out
