"""Host-side mirror of the reference's TX operator surface: make_gr_mod_4fsk / make_gr_mod_qpsk
(/root/reference/src/gr/gr_mod_4fsk.h:46-48, gr_mod_qpsk.h) batched over channels.  gr_byte_source
(src/gr/gr_byte_source.cpp:54-106) hands the modulator a byte vector; `work(bytes)` is that hand-off."""
import ctypes as C

import numpy as np

from .lib import KIND, PARAM, QrlError, check, load_library


class TxBlock:
    def __init__(self, kind, sps, samp_rate, carrier_freq, filter_width, flag=0, n_channels=1,
                 max_items=1 << 12, device=0):
        self._L = load_library()
        self.n_channels, self.max_items = int(n_channels), int(max_items)
        self._h = C.c_void_p()
        rc = self._L.qrl_tx_create(kind, sps, samp_rate, carrier_freq, filter_width, int(flag),
                                   self.n_channels, self.max_items, device, C.byref(self._h))
        if rc != 0:
            raise QrlError("qrl_tx_create failed (%d): %s" % (rc, (self._L.qrl_last_error(None) or b"").decode()))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.qrl_tx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def zero_samples(self, byte_offset, n_samples, channel=-1):
        """gr_mod_dmr: the "zero_samples" tag on input byte `byte_offset` of `channel` (-1: all), see qrl_tx_zero_samples."""
        check(self._L.qrl_tx_zero_samples(self._h, int(channel), int(byte_offset), int(n_samples)), self._h, "zero_samples")

    def set_filter_width(self, filter_width):
        """gr_mod_nbfm::set_filter_width (gr_mod_nbfm.cpp:78-93), mid-stream."""
        check(self._L.qrl_tx_set_param(self._h, -1, PARAM.FILTER_WIDTH, float(filter_width)), self._h, "set_filter_width")

    def set_ctcss(self, value):
        """gr_mod_nbfm::set_ctcss (gr_mod_nbfm.cpp:101-139): 0 = no tone (audio gain 0.98), f = CTCSS tone at f Hz (gain 0.85, 300 Hz high-pass)."""
        check(self._L.qrl_tx_set_param(self._h, -1, PARAM.CTCSS, float(value)), self._h, "set_ctcss")

    def set_bb_gain(self, value):
        check(self._L.qrl_tx_set_param(self._h, -1, PARAM.BB_GAIN, float(value)), self._h, "set_bb_gain")

    def set_stream(self, cuda_stream_ptr):
        check(self._L.qrl_tx_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), self._h, "qrl_tx_set_stream")

    def work(self, data):
        """data: uint8 [n_channels, n] frame bytes -> complex64 [n_channels, n_out] at 1 Msps."""
        data = np.ascontiguousarray(data, np.uint8)
        if data.ndim == 1:
            data = data[None, :]
        n = data.shape[1]
        check(self._L.qrl_tx_work(self._h, data.ctypes.data_as(C.c_void_p), n, n, 0), self._h, "qrl_tx_work")
        dptr, stride, nout = C.c_void_p(), C.c_long(), C.c_long()
        check(self._L.qrl_tx_out_device(self._h, C.byref(dptr), C.byref(stride), C.byref(nout)), self._h, "tx_out_device")
        out = np.zeros((self.n_channels, nout.value), np.complex64)
        n2 = C.c_long()
        check(self._L.qrl_tx_read(self._h, out.ctypes.data_as(C.c_void_p), nout.value, C.byref(n2), 0), self._h, "qrl_tx_read")
        return out[:, :n2.value]

    def work_audio(self, audio):
        """analog modulators (NBFM / SSB): float32 [n_channels, n] audio at 8 ksps -> complex64 [n_channels, n_out]."""
        audio = np.ascontiguousarray(audio, np.float32)
        if audio.ndim == 1:
            audio = audio[None, :]
        n = audio.shape[1]
        check(self._L.qrl_tx_work(self._h, audio.ctypes.data_as(C.c_void_p), n, n, 0), self._h, "qrl_tx_work")
        dptr, stride, nout = C.c_void_p(), C.c_long(), C.c_long()
        check(self._L.qrl_tx_out_device(self._h, C.byref(dptr), C.byref(stride), C.byref(nout)), self._h, "tx_out_device")
        out = np.zeros((self.n_channels, max(1, nout.value)), np.complex64)
        n2 = C.c_long()
        check(self._L.qrl_tx_read(self._h, out.ctypes.data_as(C.c_void_p), max(1, nout.value), C.byref(n2), 0), self._h, "qrl_tx_read")
        return out[:, :n2.value]

    def work_device(self, dev_ptr, n, stride):
        check(self._L.qrl_tx_work(self._h, C.c_void_p(dev_ptr), n, stride, 1), self._h, "qrl_tx_work")

    def sync(self):
        check(self._L.qrl_tx_sync(self._h), self._h, "qrl_tx_sync")

    @property
    def launches(self):
        return self._L.qrl_tx_launch_count(self._h)


def make_gr_mod_4fsk(sps, samp_rate, carrier_freq, filter_width, fm, n_channels=1, **kw):
    return TxBlock(KIND.MOD_4FSK, sps, samp_rate, carrier_freq, filter_width, int(bool(fm)), n_channels, **kw)


def make_gr_mod_qpsk(sps, samp_rate, carrier_freq, filter_width, n_channels=1, **kw):
    return TxBlock(KIND.MOD_QPSK, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_bpsk(sps, samp_rate, carrier_freq, filter_width, n_channels=1, **kw):
    """src/gr/gr_mod_bpsk.h (instances gr_mod_base.cpp:168-169)."""
    return TxBlock(KIND.MOD_BPSK, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_2fsk(sps, samp_rate, carrier_freq, filter_width, fm, n_channels=1, **kw):
    """src/gr/gr_mod_2fsk.h (instances gr_mod_base.cpp:155-159)."""
    return TxBlock(KIND.MOD_2FSK, sps, samp_rate, carrier_freq, filter_width, int(bool(fm)), n_channels, **kw)


def make_gr_mod_gmsk(sps, samp_rate, carrier_freq, filter_width, n_channels=1, **kw):
    """src/gr/gr_mod_gmsk.h (instances gr_mod_base.cpp:160-162: sps 50 / 100 / 10 = GMSK2K / 1K / 10K)."""
    return TxBlock(KIND.MOD_GMSK, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_nbfm(sps, samp_rate, carrier_freq, filter_width, n_channels=1, **kw):
    """src/gr/gr_mod_nbfm.h (instances gr_mod_base.cpp:171-172); feed with TxBlock.work_audio."""
    return TxBlock(KIND.MOD_NBFM, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_ssb(sps, samp_rate, carrier_freq, filter_width, sb, n_channels=1, **kw):
    """src/gr/gr_mod_ssb.h (instances gr_mod_base.cpp:178-179); feed with TxBlock.work_audio."""
    return TxBlock(KIND.MOD_SSB, sps, samp_rate, carrier_freq, filter_width, int(sb), n_channels, **kw)


def make_gr_mod_m17(sps=125, samp_rate=1000000, carrier_freq=1700, filter_width=9000, n_channels=1, **kw):
    """src/gr/gr_mod_m17.h:43-44 (defaults as there); items: frame bytes, 4 symbols per byte, 125 / 3 output samples per 24 ksps
    sample."""
    return TxBlock(KIND.MOD_M17, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_am(sps=125, samp_rate=1000000, carrier_freq=1700, filter_width=5000, n_channels=1, **kw):
    """src/gr/gr_mod_am.h (instance gr_mod_base.cpp:167: make_gr_mod_am(125, 1e6, 1700, 5000)); feed with TxBlock.work_audio."""
    return TxBlock(KIND.MOD_AM, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)


def make_gr_mod_dsss(sps=25, samp_rate=1000000, carrier_freq=1700, filter_width=200, n_channels=1, max_items=8, **kw):
    """src/gr/gr_mod_dsss.h (instance gr_mod_base.cpp:170: make_gr_mod_dsss(25, 1e6, 1700, 200)); items: frame bytes at 1 byte/s: every byte
    becomes 10^6 output samples, hence the small default max_items."""
    return TxBlock(KIND.MOD_DSSS, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, max_items=max_items, **kw)


def make_gr_mod_dmr(sps=125, samp_rate=1000000, carrier_freq=1700, filter_width=5000, n_channels=1, **kw):
    """src/gr/gr_mod_dmr.h:37-38 (defaults as there); items: frame bytes, 4 symbols per byte.  TxBlock.zero_samples is the
    "zero_samples" stream tag gr_dmr_source attaches to the bytes of an idle burst (gr_zero_idle_bursts)."""
    return TxBlock(KIND.MOD_DMR, sps, samp_rate, carrier_freq, filter_width, 0, n_channels, **kw)
