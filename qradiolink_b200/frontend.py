"""Front end at device rates >= 2 Msps: host-side mirror of gr_demod_base::set_samp_rate / set_carrier_offset for that case
(/root/reference/src/gr/gr_demod_base.cpp:1220-1225, 1303-1362): rotator at the device rate + rational_resampler_ccf(1, samp_rate / 1e6)
for a batch of channels.  The output stays in HBM in the layout RxBlock.work_device takes."""
import ctypes as C

import numpy as np

from .lib import QrlError, check, load_library


class Frontend:
    def __init__(self, samp_rate, n_channels=1, max_in=1 << 21, device=0):
        self._L = load_library()
        self.n_channels, self.max_in, self.decimation = int(n_channels), int(max_in), int(samp_rate) // 1000000
        self._h = C.c_void_p()
        rc = self._L.qrl_frontend_create(int(samp_rate), self.n_channels, self.max_in, device, C.byref(self._h))
        if rc != 0:
            raise QrlError("qrl_frontend_create failed (%d): %s" % (rc, (self._L.qrl_last_error(None) or b"").decode()))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.qrl_frontend_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        check(self._L.qrl_frontend_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), self._h, "qrl_frontend_set_stream")

    def set_carrier_offset(self, hz, channel=-1):
        check(self._L.qrl_frontend_set_carrier_offset(self._h, channel, float(hz)), self._h, "qrl_frontend_set_carrier_offset")

    def work(self, iq):
        """iq: complex64 [n_channels, T] at the device rate (host) -> complex64 [n_channels, n_out] at 1 Msps."""
        iq = np.ascontiguousarray(iq, np.complex64)
        if iq.ndim == 1:
            iq = iq[None, :]
        T = iq.shape[1]
        n = C.c_long()
        check(self._L.qrl_frontend_work(self._h, iq.ctypes.data_as(C.c_void_p), T, T, 0, C.byref(n)), self._h, "qrl_frontend_work")
        out = np.zeros((self.n_channels, max(1, n.value)), np.complex64)
        check(self._L.qrl_frontend_read(self._h, out.ctypes.data_as(C.c_void_p), out.shape[1]), self._h, "qrl_frontend_read")
        return out[:, :n.value]

    def work_device(self, dev_ptr, T, stride):
        n = C.c_long()
        check(self._L.qrl_frontend_work(self._h, C.c_void_p(dev_ptr), T, stride, 1, C.byref(n)), self._h, "qrl_frontend_work")
        return n.value

    def out_device(self):
        """(device pointer, stride, items) of the last call's 1 Msps output: pass to RxBlock.work_device(ptr, items, stride)."""
        p, s, n = C.c_void_p(), C.c_long(), C.c_long()
        check(self._L.qrl_frontend_out_device(self._h, C.byref(p), C.byref(s), C.byref(n)), self._h, "qrl_frontend_out_device")
        return p.value, s.value, n.value
