"""ctypes binding of the CPU ORACLE (oracle/libqrl_oracle.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package qradiolink_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# hier-block kinds (mirror qrl_oracle.h)
DEMOD_NBFM, DEMOD_4FSK, DEMOD_QPSK, DEMOD_BPSK, DEMOD_2FSK, DEMOD_SSB, DEMOD_AM, DEMOD_GMSK, DEMOD_WBFM, DEMOD_M17, DEMOD_DMR, DEMOD_DSSS = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12
MOD_4FSK, MOD_QPSK, MOD_NBFM, MOD_BPSK, MOD_2FSK, MOD_SSB, MOD_GMSK, MOD_M17, MOD_DMR, MOD_DSSS, MOD_AM = 101, 102, 103, 104, 105, 106, 107, 108, 109, 110, 111
WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECT, WIN_KAISER, WIN_BLACKMAN_HARRIS = 0, 1, 2, 3, 4, 5


def build(force=False):
    so = os.path.join(_HERE, "libqrl_oracle.so")
    src = os.path.join(_HERE, "qrl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libqrl_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, f32p, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.qo_rx_create.restype = vp
        L.qo_rx_create.argtypes = [C.c_int] * 6
        L.qo_rx_destroy.argtypes = [vp]
        L.qo_rx_work.argtypes = [vp, vp, C.c_long]
        L.qo_rx_set_carrier_offset.argtypes = [vp, C.c_double, C.c_double]
        L.qo_rx_set_param.argtypes = [vp, C.c_int, C.c_double]
        L.qo_rx_port_items.restype = C.c_long
        L.qo_rx_port_items.argtypes = [vp, C.c_int]
        L.qo_rx_port_data.restype = vp
        L.qo_rx_port_data.argtypes = [vp, C.c_int]
        L.qo_rx_port_clear.argtypes = [vp, C.c_int]
        L.qo_rx_dbg_items.restype = C.c_long
        L.qo_rx_dbg_items.argtypes = [vp, C.c_char_p]
        L.qo_rx_dbg_data.restype = vp
        L.qo_rx_dbg_data.argtypes = [vp, C.c_char_p]
        L.qo_rx_ntaps.argtypes = [vp, C.c_int, vp, C.c_int]
        L.qo_tx_create.restype = vp
        L.qo_tx_create.argtypes = [C.c_int] * 6
        L.qo_tx_destroy.argtypes = [vp]
        L.qo_tx_set_bb_gain.argtypes = [vp, C.c_float]
        L.qo_tx_zero_samples.argtypes = [vp, C.c_longlong, C.c_long]
        L.qo_tx_set_param.argtypes = [vp, C.c_int, C.c_double]
        L.qo_dsss_decoder_taps.argtypes = [vp, C.c_int, C.c_int, vp]
        L.qo_dsss_decoder_run.restype = C.c_long
        L.qo_dsss_decoder_run.argtypes = [vp, C.c_int, C.c_int, vp, C.c_long, C.c_long, vp, C.c_long]
        L.qo_mmdvm_rx_create.restype = vp
        L.qo_mmdvm_rx_create.argtypes = [C.c_int]
        L.qo_mmdvm_rx_destroy.argtypes = [vp]
        L.qo_mmdvm_rx_create2.restype = vp
        L.qo_mmdvm_rx_create2.argtypes = [C.c_int, C.c_int]
        L.qo_mmdvm_tx_create2.restype = vp
        L.qo_mmdvm_tx_create2.argtypes = [C.c_int, C.c_int]
        L.qo_mmdvm_tx_set_bb_gain.argtypes = [vp, C.c_float]
        L.qo_mmdvm_tx_zero_samples.argtypes = [vp, C.c_longlong, C.c_long]
        L.qo_mmdvm_rx_calibrate_rssi.argtypes = [vp, C.c_float]
        L.qo_mmdvm_rx_work.argtypes = [vp, vp, C.c_long]
        L.qo_mmdvm_rx_out_items.restype = C.c_long
        L.qo_mmdvm_rx_out_items.argtypes = [vp]
        L.qo_mmdvm_rx_out_data.restype = vp
        L.qo_mmdvm_rx_out_data.argtypes = [vp]
        L.qo_mmdvm_rx_rssi_items.restype = C.c_long
        L.qo_mmdvm_rx_rssi_items.argtypes = [vp]
        L.qo_mmdvm_rx_rssi_db.restype = vp
        L.qo_mmdvm_rx_rssi_db.argtypes = [vp]
        L.qo_mmdvm_rx_rssi_at.restype = vp
        L.qo_mmdvm_rx_rssi_at.argtypes = [vp]
        L.qo_mmdvm_rx_clear.argtypes = [vp]
        L.qo_rssi_tags_run.restype = C.c_long
        L.qo_rssi_tags_run.argtypes = [vp, C.c_long, C.c_float, vp, vp, C.c_long]
        L.qo_mmdvm_tx_create.restype = vp
        L.qo_mmdvm_tx_create.argtypes = [C.c_int]
        L.qo_mmdvm_tx_destroy.argtypes = [vp]
        L.qo_mmdvm_tx_work.argtypes = [vp, vp, C.c_long]
        L.qo_mmdvm_tx_out_items.restype = C.c_long
        L.qo_mmdvm_tx_out_items.argtypes = [vp]
        L.qo_mmdvm_tx_out_data.restype = vp
        L.qo_mmdvm_tx_out_data.argtypes = [vp]
        L.qo_mmdvm_tx_clear.argtypes = [vp]
        L.qo_spectrum_create.restype = vp
        L.qo_spectrum_create.argtypes = [C.c_int, C.c_int]
        L.qo_spectrum_destroy.argtypes = [vp]
        L.qo_spectrum_set_enabled.argtypes = [vp, C.c_int]
        L.qo_spectrum_set_fft_size.argtypes = [vp, C.c_int]
        L.qo_spectrum_work.argtypes = [vp, vp, C.c_long]
        L.qo_spectrum_get.argtypes = [vp, vp]
        L.qo_zero_idle_run.argtypes = [vp, C.c_long, C.c_uint, vp, vp, C.c_long, vp]
        L.qo_zero_idle_run.restype = None
        L.qo_tx_work.argtypes = [vp, vp, C.c_long]
        L.qo_tx_out_items.restype = C.c_long
        L.qo_tx_out_items.argtypes = [vp]
        L.qo_tx_out_data.restype = vp
        L.qo_tx_out_data.argtypes = [vp]
        L.qo_tx_out_clear.argtypes = [vp]
        for name in ("qo_firdes_low_pass",):
            getattr(L, name).argtypes = [C.c_double] * 4 + [C.c_int, vp, C.c_int]
        L.qo_firdes_low_pass_2.argtypes = [C.c_double] * 5 + [C.c_int, vp, C.c_int]
        L.qo_firdes_band_pass.argtypes = [C.c_double] * 5 + [C.c_int, vp, C.c_int]
        L.qo_firdes_band_pass_2.argtypes = [C.c_double] * 6 + [C.c_int, vp, C.c_int]
        L.qo_firdes_complex_band_pass.argtypes = [C.c_double] * 5 + [C.c_int, vp, C.c_int]
        L.qo_firdes_complex_band_pass_2.argtypes = [C.c_double] * 6 + [C.c_int, vp, C.c_int]
        L.qo_firdes_rrc.argtypes = [C.c_double] * 4 + [C.c_int, vp, C.c_int]
        L.qo_firdes_gaussian.argtypes = [C.c_double] * 3 + [C.c_int, vp, C.c_int]
        L.qo_deemph_taps.argtypes = [C.c_int, C.c_double, vp, vp]
        L.qo_preemph_taps.argtypes = [C.c_int, C.c_double, C.c_double, vp, vp]
        L.qo_sincosf.argtypes = [C.c_float, vp, vp]
        L.qo_fast_atan2f.restype = C.c_float
        L.qo_fast_atan2f.argtypes = [C.c_float, C.c_float]
        L.qo_clock_loop_gains.argtypes = [C.c_float] * 3 + [vp, vp]
        L.qo_control_loop_gains.argtypes = [C.c_float, vp, vp]
        L.qo_fir_decim_ccf.restype = C.c_long
        L.qo_fir_decim_ccf.argtypes = [vp, C.c_int, C.c_int, vp, C.c_long, vp]
        L.qo_fir_fff.restype = C.c_long
        L.qo_fir_fff.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp, C.c_long]
        for name in ("qo_cc_encode", "qo_cc_decode"):
            getattr(L, name).restype = C.c_long
            getattr(L, name).argtypes = [vp, C.c_long, vp]
        L.qo_scramble.argtypes = [vp, C.c_long, vp]
        L.qo_descramble.argtypes = [vp, C.c_long, vp]
        L.qo_find_frames.restype = C.c_long
        L.qo_find_frames.argtypes = [vp, C.c_long, C.c_uint32, C.c_int, C.c_int, vp, C.c_long]
        L.qo_rssi_create.restype = vp
        L.qo_rssi_create.argtypes = [C.c_float]
        L.qo_rssi_destroy.argtypes = [vp]
        L.qo_rssi_work.restype = C.c_float
        L.qo_rssi_work.argtypes = [vp, vp, C.c_long]
        L.qo_deframer_create.restype = vp
        L.qo_deframer_create.argtypes = [C.c_int] * 3
        L.qo_deframer_destroy.argtypes = [vp]
        L.qo_deframer_work.restype = C.c_long
        L.qo_deframer_work.argtypes = [vp, vp, C.c_long, vp, C.c_int, C.c_long]
        L.qo_deframer_modem_sync.argtypes = [vp]
        L.qo_frontend_create.restype = vp
        L.qo_frontend_create.argtypes = [C.c_int]
        L.qo_frontend_destroy.argtypes = [vp]
        L.qo_frontend_ntaps.argtypes = [vp]
        L.qo_frontend_set_carrier_offset.argtypes = [vp, C.c_double]
        L.qo_frontend_work.restype = C.c_long
        L.qo_frontend_work.argtypes = [vp, vp, C.c_long, vp, C.c_long]
        L.qo_frame.restype = C.c_long
        L.qo_frame.argtypes = [vp, C.c_long, C.c_uint32, C.c_int, C.c_int, vp, C.c_long]
        L.qo_pfb_channelizer_create.restype = vp
        L.qo_pfb_channelizer_create.argtypes = [C.c_int, vp, C.c_int]
        L.qo_pfb_synthesizer_create.restype = vp
        L.qo_pfb_synthesizer_create.argtypes = [C.c_int, vp, C.c_int]
        L.qo_pfb_destroy.argtypes = [vp]
        L.qo_pfb_channelizer_work.restype = C.c_long
        L.qo_pfb_channelizer_work.argtypes = [vp, vp, C.c_long, vp, C.c_long, C.c_long]
        L.qo_pfb_synthesizer_work.restype = C.c_long
        L.qo_pfb_synthesizer_work.argtypes = [vp, vp, C.c_long, C.c_long, vp]
        L.qo_cessb_clipper.argtypes = [vp, C.c_long, C.c_float, vp]
        L.qo_cessb_stretcher.restype = C.c_long
        L.qo_cessb_stretcher.argtypes = [vp, C.c_long, vp]
        L.qo_disc4.argtypes = [vp, vp, vp, vp, C.c_long, vp]
        L.qo_dfbb_create.restype = vp
        L.qo_dfbb_create.argtypes = [C.c_int]
        L.qo_dfbb_destroy.argtypes = [vp]
        L.qo_dfbb_work.restype = C.c_long
        L.qo_dfbb_work.argtypes = [vp, vp, C.c_long, vp, C.c_long]
        _LIB = L
    return _LIB


_REF = None


def ref_blocks():
    """oracle/_ref/libqrl_ref_blocks.so: the reference's in-tree GNU Radio blocks compiled unmodified (oracle/Makefile `ref`),
    or None when it was never built (the reference tree exists only in the build container; the built .so travels)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libqrl_ref_blocks.so")
        if not os.path.exists(so):
            return None
        lib()                                             # libqrl_oracle.so first: the shim resolves two symbols from it
        R = C.CDLL(so)
        vp = C.c_void_p
        R.ref_block_destroy.argtypes = [vp]
        R.ref_disc4.argtypes = [vp, vp, vp, vp, C.c_long, vp]
        R.ref_cessb_clipper.restype = C.c_long
        R.ref_cessb_clipper.argtypes = [vp, C.c_long, C.c_float, vp]
        R.ref_cessb_stretcher.restype = C.c_long
        R.ref_cessb_stretcher.argtypes = [vp, C.c_long, C.c_long, vp]
        R.ref_dfbb_create.restype = vp
        R.ref_dfbb_create.argtypes = [C.c_int]
        R.ref_dfbb_work.restype = C.c_long
        R.ref_dfbb_work.argtypes = [vp, vp, C.c_long, vp, C.c_long]
        for kind, ty in (("bit", vp), ("audio", vp), ("const", vp)):
            getattr(R, "ref_%s_sink_create" % kind).restype = vp
            getattr(R, "ref_%s_sink_work" % kind).argtypes = [vp, ty, C.c_int]
            getattr(R, "ref_%s_sink_get" % kind).restype = C.c_long
            getattr(R, "ref_%s_sink_get" % kind).argtypes = [vp, ty, C.c_long]
        R.ref_dsss_encode.restype = C.c_long
        R.ref_dsss_encode.argtypes = [vp, C.c_int, vp, C.c_long, vp]
        R.ref_dsss_decoder_create.restype = vp
        R.ref_dsss_decoder_create.argtypes = [vp, C.c_int, C.c_float]
        R.ref_dsss_decoder_taps.argtypes = [vp, vp, C.c_int]
        R.ref_dsss_decoder_history.argtypes = [vp]
        R.ref_dsss_decoder_work.restype = C.c_long
        R.ref_dsss_decoder_work.argtypes = [vp, vp, C.c_int, vp, vp]
        R.ref_zero_idle.restype = C.c_long
        R.ref_zero_idle.argtypes = [vp, C.c_long, C.c_uint, vp, vp, C.c_long, vp, C.c_long, vp]
        R.ref_rx_fft_create.restype = vp
        R.ref_rx_fft_create.argtypes = [C.c_uint, C.c_int]
        R.ref_rx_fft_set_enabled.argtypes = [vp, C.c_int]
        R.ref_rx_fft_set_fft_size.argtypes = [vp, C.c_uint]
        R.ref_rx_fft_work.argtypes = [vp, vp, C.c_int]
        R.ref_rx_fft_get.restype = C.c_uint
        R.ref_rx_fft_get.argtypes = [vp, vp]
        R.ref_sample_sink_create.restype = vp
        R.ref_sample_sink_set_enabled.argtypes = [vp, C.c_int]
        R.ref_sample_sink_set_window.argtypes = [vp, C.c_uint]
        R.ref_sample_sink_work.argtypes = [vp, vp, C.c_int]
        R.ref_sample_sink_get.restype = C.c_long
        R.ref_sample_sink_get.argtypes = [vp, vp, C.c_long]
        R.ref_rssi_tags.restype = C.c_long
        R.ref_rssi_tags.argtypes = [vp, C.c_long, C.c_float, vp, C.c_long, vp, vp, C.c_long]
        _REF = R
    return _REF


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ design
def _taps(fn, *args, complex_out=False, cap=1 << 15):
    out = np.zeros(cap * (2 if complex_out else 1), np.float32)
    n = fn(*args, _p(out), cap)
    assert n > 0, n
    if complex_out:
        return out[: 2 * n].view(np.complex64).copy()
    return out[:n].copy()


def low_pass(gain, fs, fc, tw, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_low_pass, gain, fs, fc, tw, win)


def low_pass_2(gain, fs, fc, tw, att, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_low_pass_2, gain, fs, fc, tw, att, win)


def band_pass_2(gain, fs, lo, hi, tw, att, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_band_pass_2, gain, fs, lo, hi, tw, att, win)


def complex_band_pass_2(gain, fs, lo, hi, tw, att, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_complex_band_pass_2, gain, fs, lo, hi, tw, att, win, complex_out=True)


def band_pass(gain, fs, lo, hi, tw, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_band_pass, gain, fs, lo, hi, tw, win)


def complex_band_pass(gain, fs, lo, hi, tw, win=WIN_HAMMING):
    return _taps(lib().qo_firdes_complex_band_pass, gain, fs, lo, hi, tw, win, complex_out=True)


def rrc(gain, fs, symrate, alpha, ntaps):
    return _taps(lib().qo_firdes_rrc, gain, fs, symrate, alpha, ntaps)


def gaussian(gain, spb, bt, ntaps):
    return _taps(lib().qo_firdes_gaussian, gain, spb, bt, ntaps)


def deemph_taps(fs, tau):
    a = np.zeros(2); b = np.zeros(2)
    lib().qo_deemph_taps(fs, tau, _p(a), _p(b))
    return a, b


def preemph_taps(fs, tau, fh=-1.0):
    a = np.zeros(2); b = np.zeros(2)
    lib().qo_preemph_taps(fs, tau, fh, _p(a), _p(b))
    return a, b


def table(name):
    n = {"atan": 257, "mmse": 129 * 8, "tanh": 256, "fxpt_sine": 2048}[name]
    out = np.zeros(n, np.float32)
    getattr(lib(), "qo_%s_table" % name)(_p(out))
    return out


def sincosf(x):
    x = np.asarray(x, np.float32).ravel()
    s = np.zeros_like(x); c = np.zeros_like(x)
    sv = C.c_float(); cv = C.c_float()
    L = lib()
    for i, v in enumerate(x):
        L.qo_sincosf(float(v), C.byref(sv), C.byref(cv))
        s[i] = sv.value; c[i] = cv.value
    return s, c


def fir_decim_ccf(h, D, x):
    h = np.ascontiguousarray(h, np.float32); x = np.ascontiguousarray(x, np.complex64)
    y = np.zeros(len(x) // D + 2, np.complex64)
    n = lib().qo_fir_decim_ccf(_p(h), len(h), D, _p(x), len(x), _p(y))
    return y[:n]


def fir_fff(h, L_, M, x):
    h = np.ascontiguousarray(h, np.float32); x = np.ascontiguousarray(x, np.float32)
    cap = len(x) * L_ // M + 8
    y = np.zeros(cap, np.float32)
    n = lib().qo_fir_fff(_p(h), len(h), L_, M, _p(x), len(x), _p(y), cap)
    return y[:n]


def cc_encode(bits):
    bits = np.ascontiguousarray(bits, np.uint8); out = np.zeros(2 * len(bits), np.uint8)
    n = lib().qo_cc_encode(_p(bits), len(bits), _p(out)); return out[:n]


def cc_decode(soft):
    soft = np.ascontiguousarray(soft, np.uint8); out = np.zeros(len(soft) // 2 + 80, np.uint8)
    n = lib().qo_cc_decode(_p(soft), len(soft), _p(out)); return out[:n]


def scramble(bits):
    bits = np.ascontiguousarray(bits, np.uint8); out = np.zeros_like(bits)
    lib().qo_scramble(_p(bits), len(bits), _p(out)); return out


def descramble(bits):
    bits = np.ascontiguousarray(bits, np.uint8); out = np.zeros_like(bits)
    lib().qo_descramble(_p(bits), len(bits), _p(out)); return out


def find_frames(bits, sync, sync_bits, frame_len):
    bits = np.ascontiguousarray(bits, np.uint8)
    maxf = len(bits) // (8 * frame_len) + 1
    out = np.zeros(maxf * frame_len, np.uint8)
    n = lib().qo_find_frames(_p(bits), len(bits), sync, sync_bits, frame_len, _p(out), maxf)
    return out[: n * frame_len].reshape(n, frame_len)


# ------------------------------------------------------------------ chains
class Rx:
    """One channel of a reference demod hier-block (make_gr_demod_*), CPU oracle."""
    PORT_DTYPES = {0: np.complex64, 2: np.uint8, 3: np.uint8}   # port 1: complex constellation or float audio

    def __init__(self, kind, sps, samp_rate, carrier_freq, filter_width, flag=0):
        self.kind = kind
        self.h = lib().qo_rx_create(kind, sps, samp_rate, carrier_freq, filter_width, int(flag))
        if not self.h:
            raise ValueError("oracle: unsupported demod kind %r" % kind)

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_rx_destroy(self.h); self.h = None

    def work(self, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        rc = lib().qo_rx_work(self.h, _p(iq), len(iq))
        assert rc == 0

    def set_param(self, key, value):
        """analog-block setters (keys = qradiolink_b200.PARAM); raises for a block that has no such setter"""
        if lib().qo_rx_set_param(self.h, int(key), float(value)) != 0:
            raise ValueError("oracle: no setter %r for this block" % key)

    def set_carrier_offset(self, hz, samp_rate=1e6):
        lib().qo_rx_set_carrier_offset(self.h, float(hz), float(samp_rate))

    def port(self, p, clear=True):
        n = lib().qo_rx_port_items(self.h, p)
        dt = self.PORT_DTYPES.get(p, np.float32 if self.kind in (DEMOD_NBFM, DEMOD_SSB, DEMOD_AM, DEMOD_WBFM) else np.complex64)
        if self.kind == DEMOD_DMR and p == 3:
            dt = np.float32                                   # gr_demod_dmr port 3: the symbol filter output
        nbytes = n * np.dtype(dt).itemsize
        if n == 0:
            return np.zeros(0, dt)
        buf = C.string_at(lib().qo_rx_port_data(self.h, p), nbytes)
        out = np.frombuffer(buf, dt).copy()
        if clear:
            lib().qo_rx_port_clear(self.h, p)
        return out

    def dbg(self, name, dtype):
        n = lib().qo_rx_dbg_items(self.h, name.encode())
        if n <= 0:
            return np.zeros(0, dtype)
        buf = C.string_at(lib().qo_rx_dbg_data(self.h, name.encode()), n * np.dtype(dtype).itemsize)
        return np.frombuffer(buf, dtype).copy()

    def taps(self, which):
        out = np.zeros(4096, np.float32)
        n = lib().qo_rx_ntaps(self.h, which, _p(out), 4096)
        return out[:n].copy()


class Tx:
    """One channel of a reference mod hier-block (make_gr_mod_*), CPU oracle."""

    def __init__(self, kind, sps, samp_rate, carrier_freq, filter_width, flag=0):
        self.h = lib().qo_tx_create(kind, sps, samp_rate, carrier_freq, filter_width, int(flag))
        if not self.h:
            raise ValueError("oracle: unsupported mod kind %r" % kind)

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_tx_destroy(self.h); self.h = None

    def set_bb_gain(self, g):
        lib().qo_tx_set_bb_gain(self.h, g)

    def set_param(self, key, value):
        rc = lib().qo_tx_set_param(self.h, int(key), float(value))
        if rc != 0:
            raise ValueError("oracle: the reference has no such setter on this modulator")

    def zero_samples(self, byte_offset, n_samples):
        """MOD_DMR: the "zero_samples" stream tag on input byte `byte_offset` (gr_dmr_source.cpp:148)."""
        return lib().qo_tx_zero_samples(self.h, int(byte_offset), int(n_samples))

    def work(self, data):
        data = np.ascontiguousarray(data)
        rc = lib().qo_tx_work(self.h, _p(data), len(data))
        assert rc == 0
        n = lib().qo_tx_out_items(self.h)
        buf = C.string_at(lib().qo_tx_out_data(self.h), n * 8)
        lib().qo_tx_out_clear(self.h)
        return np.frombuffer(buf, np.complex64).copy()


def zero_idle(x, delay, tag_items, tag_vals):
    """gr_zero_idle_bursts alone (restatement): x complex64 [n], tags on the block's own input items."""
    x = np.ascontiguousarray(x, np.complex64)
    to = np.ascontiguousarray(tag_items, np.int64); tv = np.ascontiguousarray(tag_vals, np.int64)
    out = np.empty_like(x)
    lib().qo_zero_idle_run(_p(x), len(x), int(delay), _p(to), _p(tv), len(to), _p(out))
    return out


class MmdvmRx:
    """One channel of gr_demod_mmdvm_multi2 behind the channelizer (gr_demod_mmdvm_multi2.cpp:56-126): complex at 25 ksps -> int16
    discriminator samples at 24 ksps + the RSSI tags."""

    def __init__(self, filter_width=5000, single=False):
        """single=True: gr_demod_mmdvm (gr_demod_mmdvm.cpp:30-64): complex at 250 ksps in, x12 / 125, RSSI in front of the filter."""
        self.h = lib().qo_mmdvm_rx_create2(int(filter_width), int(bool(single)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_mmdvm_rx_destroy(self.h); self.h = None

    def calibrate_rssi(self, level):
        lib().qo_mmdvm_rx_calibrate_rssi(self.h, float(level))

    def work(self, x):
        """-> (int16 samples, rssi dB values, their item offsets at 24 ksps)"""
        x = np.ascontiguousarray(x, np.complex64)
        lib().qo_mmdvm_rx_work(self.h, _p(x), len(x))
        n, k = lib().qo_mmdvm_rx_out_items(self.h), lib().qo_mmdvm_rx_rssi_items(self.h)
        out = np.frombuffer(C.string_at(lib().qo_mmdvm_rx_out_data(self.h), 2 * n), np.int16).copy()
        db = np.frombuffer(C.string_at(lib().qo_mmdvm_rx_rssi_db(self.h), 4 * k), np.float32).copy()
        at = np.frombuffer(C.string_at(lib().qo_mmdvm_rx_rssi_at(self.h), 8 * k), np.int64).copy()
        lib().qo_mmdvm_rx_clear(self.h)
        return out, db, at


class MmdvmTx:
    """One channel of gr_mod_mmdvm_multi2 in front of the synthesizer (gr_mod_mmdvm_multi2.cpp:47-126): int16 at 24 ksps -> complex
    at 25 ksps."""

    def __init__(self, filter_width=5000, single=False):
        """single=True: gr_mod_mmdvm (gr_mod_mmdvm.cpp:28-70): x125 / 12 to 250 ksps, bb_gain in front of the resampler."""
        self.h = lib().qo_mmdvm_tx_create2(int(filter_width), int(bool(single)))

    def set_bb_gain(self, g):
        lib().qo_mmdvm_tx_set_bb_gain(self.h, float(g))

    def zero_samples(self, item_offset, n_samples):
        """the "zero_samples" tag on item `item_offset` of gr_zero_idle_bursts(0)'s own stream (qo_mmdvm_tx_zero_samples)"""
        return lib().qo_mmdvm_tx_zero_samples(self.h, int(item_offset), int(n_samples))

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_mmdvm_tx_destroy(self.h); self.h = None

    def work(self, s):
        s = np.ascontiguousarray(s, np.int16)
        lib().qo_mmdvm_tx_work(self.h, _p(s), len(s))
        n = lib().qo_mmdvm_tx_out_items(self.h)
        out = np.frombuffer(C.string_at(lib().qo_mmdvm_tx_out_data(self.h), 8 * n), np.complex64).copy()
        lib().qo_mmdvm_tx_clear(self.h)
        return out


class Spectrum:
    """rx_fft_c (rx_fft.cpp:44-129) restated: one work() per call; get() returns the fft-shifted dB points or None."""

    def __init__(self, fft_size=32768, window=WIN_BLACKMAN_HARRIS):
        self.h = lib().qo_spectrum_create(fft_size, window)
        if not self.h:
            raise ValueError("oracle: fft size must be a power of two")
        self.n = fft_size

    def __del__(self):
        if getattr(self, "h", None):
            lib().qo_spectrum_destroy(self.h); self.h = None

    def set_enabled(self, on):
        lib().qo_spectrum_set_enabled(self.h, int(bool(on)))

    def set_fft_size(self, n):
        lib().qo_spectrum_set_fft_size(self.h, int(n)); self.n = int(n)

    def work(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        lib().qo_spectrum_work(self.h, _p(x), len(x))

    def get(self):
        out = np.empty(self.n, np.float32)
        return out if lib().qo_spectrum_get(self.h, _p(out)) else None


class PfbChannelizer:
    """pfb_channelizer_ccf(M, taps, 1.0) behind stream_to_streams(M) (gr_demod_mmdvm_multi2.cpp:98-107), streaming."""

    def __init__(self, M, taps):
        taps = np.ascontiguousarray(taps, np.float32)
        self.M = M
        self._h = lib().qo_pfb_channelizer_create(M, _p(taps), len(taps))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_pfb_destroy(self._h); self._h = None

    def work(self, x):
        """x: complex64 [n] -> complex64 [M, n_out] (n_out = complete frames of M available)."""
        x = np.ascontiguousarray(x, np.complex64)
        cap = len(x) // self.M + 2
        out = np.zeros((self.M, cap), np.complex64)
        n = lib().qo_pfb_channelizer_work(self._h, _p(x), len(x), _p(out), cap, 0)
        return out[:, :n].copy()


class PfbSynthesizer:
    """pfb_synthesizer_ccf(M, taps, false) (gr_mod_mmdvm_multi2.cpp:90-92), streaming."""

    def __init__(self, M, taps):
        taps = np.ascontiguousarray(taps, np.float32)
        self.M = M
        self._h = lib().qo_pfb_synthesizer_create(M, _p(taps), len(taps))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_pfb_destroy(self._h); self._h = None

    def work(self, x):
        """x: complex64 [M, n] -> complex64 [n * M]."""
        x = np.ascontiguousarray(x, np.complex64)
        out = np.zeros(x.shape[1] * self.M, np.complex64)
        lib().qo_pfb_synthesizer_work(self._h, _p(x), x.shape[1], x.shape[1], _p(out))
        return out


class Deframer:
    """gr_modem::synchronize / findSync / packBytes (gr_modem.cpp:1119-1282), one channel, streaming.
    work(bits) -> list of (frame_type, payload bytes)."""

    def __init__(self, sync_class, bit_buf_len, rx_frame_length):
        self.rec_bytes = 8 + (bit_buf_len + 7) // 8 + 8
        self._h = lib().qo_deframer_create(sync_class, bit_buf_len, rx_frame_length)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_deframer_destroy(self._h); self._h = None

    def work(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        cap = len(bits) // 8 + 2
        rec = np.zeros((cap, self.rec_bytes), np.uint8)
        n = lib().qo_deframer_work(self._h, _p(bits), len(bits), _p(rec), self.rec_bytes, cap)
        out = []
        for r in rec[:n]:
            ty, nb = int(r[:4].view(np.uint32)[0]), int(r[4:8].view(np.uint32)[0])
            out.append((ty, r[8:8 + nb].tobytes()))
        return out

    @property
    def modem_sync(self):
        return lib().qo_deframer_modem_sync(self._h)


class Frontend:
    """gr_demod_base front end at a device rate >= 2 Msps (gr_demod_base.cpp:1303-1362): rotator + /N decimator to 1 Msps, one channel."""

    def __init__(self, samp_rate):
        self._h = lib().qo_frontend_create(int(samp_rate))
        if not self._h:
            raise ValueError("front end: samp_rate must be a multiple of 1e6, >= 2e6")
        self.D = int(samp_rate) // 1000000

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_frontend_destroy(self._h); self._h = None

    @property
    def ntaps(self):
        return lib().qo_frontend_ntaps(self._h)

    def set_carrier_offset(self, hz):
        lib().qo_frontend_set_carrier_offset(self._h, float(hz))

    def work(self, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        cap = len(iq) // self.D + 2
        out = np.zeros(cap, np.complex64)
        n = lib().qo_frontend_work(self._h, _p(iq), len(iq), _p(out), cap)
        return out[:n].copy()


def frame(payload, frame_type, one_k_mode=False, burst_ip=False):
    """gr_modem::frame (gr_modem.cpp:904-961): bytes that go to the modulator for one frame"""
    payload = np.ascontiguousarray(np.frombuffer(bytes(payload), np.uint8))
    out = np.zeros(len(payload) + 16, np.uint8)
    n = lib().qo_frame(_p(payload), len(payload), int(frame_type), int(one_k_mode), int(burst_ip), _p(out), len(out))
    return out[:n].copy()


class DeframerBB:
    """gr_deframer_bb (gr_deframer_bb.cpp:83-185), one channel, streaming: work(bits) -> the bits it hands to get_data()."""

    def __init__(self, modem_type):
        self._h = lib().qo_dfbb_create(modem_type)
        if not self._h:
            raise ValueError("gr_deframer_bb: modem_type must be 1, 2 or 3")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_dfbb_destroy(self._h); self._h = None

    def work(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        cap = len(bits) + len(bits) // 8 * 3 + 64
        out = np.zeros(cap, np.uint8)
        n = lib().qo_dfbb_work(self._h, _p(bits), len(bits), _p(out), cap)
        return out[:n].copy()


class Rssi:
    """rssi_block.cpp:25-45 on one channel's port-0 stream; work(iq) -> dB value probe_signal_f would hold."""

    def __init__(self, level=0.0):
        self._h = lib().qo_rssi_create(level)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_rssi_destroy(self._h); self._h = None

    def work(self, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        return float(lib().qo_rssi_work(self._h, _p(iq), len(iq)))
