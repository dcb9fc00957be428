"""CPU tier: the CUDA kernels of the two chains that have not run on a GPU yet (DMR receive, M17 modulator), compiled FROM THEIR
SOURCE in qradiolink_b200/csrc/qrl_kernels.cuh for host threads (tools/emu/: one OS thread per CUDA thread, mbarrier / cp.async.bulk
stand-ins) and compared with the oracle bit for bit.  This checks kernel arithmetic, ring indexing, the warp-specialised symbol-sync
hand-off and the launch geometry / output-count formulas the harnesses copy from qrl_b200.cu; it does NOT replace the GPU tier (the
host wiring in qrl_b200.cu and everything timing / memory-model related only shows on hardware).  Test infrastructure, like oracle/."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(900)      # pytest-timeout: a scheduling bug in the emulator must fail, not hang the tier

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "emu")
sys.path.insert(0, EMU)

M17_TX_NAMES = ["d_sine_tab", "TxBitState", "TXM_4FSK", "tx_bits_kernel", "tx_shape_fm_kernel", "fir_ccf_ring_kernel", "scale2_ring_kernel",
                "resamp_ring_ccf_generic_kernel", "ring_to_port_f32_kernel"]
DMR_RX_NAMES = ["d_atan_tab", "d_tanh_tab", "d_mmse_tab", "qrl_sincosf", "qrl_fast_atan2f", "qrl_tanhf_lut", "qrl_clip", "qrl_soft_u8",
                "qdemod_fir_fff_kernel", "ring_to_port_f32_kernel", "SL_RECT4", "EPI_4FSK_FM", "LOOP_SYMSYNC", "LoopState", "SymSyncState",
                "symsync_stride", "SYMSYNC_TAB_FLOATS", "SymSyncParams", "qrl_slice", "qrl_costas_step", "qrl_sincosf_small", "qrl_tanhf_lut_rd",
                "qrl_costas4_snr_step", "qrl_fill_tanh_s", "qrl_phase_wrap_slow", "qrl_costas4_snr_chunk", "qrl_slice_rect4", "qrl_clip1",
                "symsync_generic_step", "symsync_kernel", "symsync_ext_epilogue_kernel"]


def build(tmp_path, harness, names):
    from extract import extract
    src = open(os.path.join(ROOT, "qradiolink_b200", "csrc", "qrl_kernels.cuh")).read()
    d = tmp_path / harness
    d.mkdir()
    (d / "kernels_extracted.inc").write_text(extract(src, names))
    exe = str(d / harness)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-Wno-unknown-pragmas", "-I", str(d), "-I", EMU,
                           "-o", exe, os.path.join(EMU, harness + ".cpp")])
    return exe, d


def test_ring_to_port_kernel_emulated(tmp_path):
    exe, _ = build(tmp_path, "ring_to_port_harness", M17_TX_NAMES)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr


def test_m17_modulator_kernels_emulated_match_oracle(tmp_path, oracle):
    """tx_bits_kernel<TXM_M17> -> tx_shape_fm_kernel -> fir_ccf_ring_kernel -> scale2_ring_kernel -> resamp_ring_ccf_generic_kernel,
    three calls with carried state, against oracle.Tx(MOD_M17)."""
    exe, d = build(tmp_path, "m17_tx_harness", M17_TX_NAMES)
    C, n = 2, 48
    data = np.random.default_rng(5400).integers(0, 256, (C, n), dtype=np.uint8)
    data.tofile(str(d / "in.bin"))
    r = subprocess.run([exe, str(d / "in.bin"), str(C), str(n), str(d / "out.bin"), "1", "20"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    nout = int(r.stdout.strip())
    got = np.fromfile(str(d / "out.bin"), np.complex64).reshape(C, nout)
    for c in range(C):
        want = oracle.Tx(oracle.MOD_M17, 125, 1000000, 1700, 9000, 0).work(data[c])
        assert nout == len(want) == n * 4 * 5 * 125 // 3
        assert np.array_equal(got[c], want), c


@pytest.mark.parametrize("C,T,cuts", [(3, 300000, (3600, 3601, 3726, 5327, 5328)), (37, 60000, (700,))])
def test_dmr_receive_kernels_emulated_match_oracle(tmp_path, oracle, C, T, cuts):
    """qdemod_fir_fff_kernel -> ring_to_port_f32_kernel / symsync_kernel<.., EPI_EXT_4FSK_FM, 256, 3, 1, LOOP_SYMSYNC, 3> (TMA producer,
    recurrence with the plain M&M detector, drain warp) -> symsync_ext_epilogue_kernel<1>, ragged calls, channel counts that leave
    lanes idle; input = the oracle's port 0, outputs = ports 1 (symbols), 2 (hard bits), 3 (symbol filter output)."""
    from tests.test_gpu_m17 import m17_like_signal
    exe, d = build(tmp_path, "dmr_rx_harness", DMR_RX_NAMES)
    rng = np.random.default_rng(91)
    X = np.stack([m17_like_signal(rng, T) for _ in range(C)])
    ref = []
    for c in range(C):
        rx = oracle.Rx(oracle.DEMOD_DMR, 5, 1000000, 0, 0, 0)
        rx.work(X[c])
        ref.append([rx.port(p) for p in range(4)])
    n = len(ref[0][0])
    np.stack([r[0] for r in ref]).tofile(str(d / "in.bin"))
    r = subprocess.run([exe, str(d / "in.bin"), str(C), str(n), str(d / "out"), *[str(v) for v in cuts if v < n]],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stdout + r.stderr
    for c in range(C):
        p1 = np.fromfile(str(d / ("out.p1.%d.bin" % c)), np.complex64)
        p2 = np.fromfile(str(d / ("out.p2.%d.bin" % c)), np.uint8)
        p3 = np.fromfile(str(d / ("out.p3.%d.bin" % c)), np.float32)
        assert len(p1) > 100 and np.array_equal(p1, ref[c][1]), (c, len(p1), len(ref[c][1]))
        assert np.array_equal(p2, ref[c][2]), c
        assert np.array_equal(p3, ref[c][3]), c
