"""Case definitions shared by tests/golden/make_golden.py (which freezes the answers) and tests/test_golden.py
(which checks the oracle and the CUDA path against the frozen answers).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

EMPHASIS_CASES = ((20000, 50e-6), (8000, 50e-6), (48000, 75e-6), (8000, 75e-6))


def load_ref_emphasis():
    so = os.path.join(ROOT, "oracle", "_ref", "libqrl_ref_emphasis.so")
    return C.CDLL(so) if os.path.exists(so) else None


def ref_deemph(R, fs, tau):
    a = np.zeros(2); b = np.zeros(2)
    R.ref_deemph_taps(C.c_int(fs), C.c_double(tau), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    return list(a) + list(b)


def ref_preemph(R, fs, tau):
    a = np.zeros(2); b = np.zeros(2)
    R.ref_preemph_taps(C.c_int(fs), C.c_double(tau), C.c_double(-1.0), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    return list(a) + list(b)


# design functions with the argument lists the reference uses (SURVEY 8a tap counts: 419, 55, 251, 837, 1045, 689, 23)
DESIGN_CASES = {
    "lp_1e6_10k_10k_bh_419": lambda O: O.low_pass(1, 1e6, 10000, 10000, O.WIN_BLACKMAN_HARRIS),
    "lp_20k_3000_1500_bh_55": lambda O: O.low_pass(1, 20000, 3000, 1500, O.WIN_BLACKMAN_HARRIS),
    "rrc_1.5_20k_2k_0.2_251": lambda O: O.rrc(1.5, 20000, 2000, 0.2, 251),
    "lp_20k_2000_100_bh_837": lambda O: O.low_pass(1, 20000, 2000, 100, O.WIN_BLACKMAN_HARRIS),
    "lp_20_1e6_3500_3500_hamming_689": lambda O: O.low_pass(20, 1e6, 3500, 3500, O.WIN_HAMMING),
    "rrc_2_2_1_0.35_23": lambda O: O.rrc(2, 2, 1, 0.35, 23),
    "lp2_1e6_250k_50k_60_bh_55": lambda O: O.low_pass_2(1, 1e6, 250000, 50000, 60, O.WIN_BLACKMAN_HARRIS),
    "table_atan": lambda O: O.table("atan"),
    "table_mmse": lambda O: O.table("mmse"),
    "table_tanh": lambda O: O.table("tanh"),
}


def _sig_4fsk(fm):
    return lambda O, sg: sg.gen_4fsk_channels(2, 1 << 18, seed0=9100 if fm else 9200, fm=fm)[0]


def _sig_qpsk(O, sg):
    return sg.gen_qpsk_channels(1, 1 << 17, seed0=9300)[0]


def _sig_nbfm(O, sg):
    return sg.gen_nbfm_channels(1, 1 << 18, seed0=9400)


def _sig_digital(kind):
    def f(O, sg):
        rng = np.random.default_rng(9500 + (1 if kind == "bpsk" else 2))
        T = 1 << 18
        data, _ = sg.frames_4fsk(rng, 4)
        if kind == "bpsk":
            iq = O.Tx(O.MOD_BPSK, 250, 1000000, 1700, 2800, 0).work(data)
        else:
            iq = O.Tx(O.MOD_2FSK, 25, 1000000, 1700, 4000, 1).work(data)
        x = sg.channel(iq, rng, fo_hz=60.0, phase=1.0, delay=100, snr_db=18.0, amp=0.1, total=T)
        return x[None, :]
    return f


def _sig_am(O, sg):
    rng = np.random.default_rng(9650)
    T = 1 << 18
    n = np.arange(T)
    aud = 0.5 * np.sin(2 * np.pi * 900.0 * n / 1e6) + 0.3 * np.sin(2 * np.pi * 2100.0 * n / 1e6)
    x = 0.4 * (1.0 + 0.8 * aud) * np.exp(2j * np.pi * 150.0 * n / 1e6)
    x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.004
    return x.astype(np.complex64)[None, :]


def _sig_wbfm(O, sg):
    rng = np.random.default_rng(9660)
    T = 1 << 17
    n = np.arange(T)
    aud = 0.6 * np.sin(2 * np.pi * 1000.0 * n / 1e6) + 0.3 * np.sin(2 * np.pi * 2900.0 * n / 1e6)
    x = 0.5 * np.exp(1j * 2 * np.pi * 50000.0 * np.cumsum(aud) / 1e6)
    x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.005
    return x.astype(np.complex64)[None, :]


def _sig_m17(O, sg):
    rng = np.random.default_rng(9670)
    T = 1 << 18
    nsym = int(T / 1e6 * 4800) + 2
    sy = np.array([-1.5, -0.5, 0.5, 1.5])[rng.integers(0, 4, nsym)]
    t = np.arange(T) / 1e6
    x = sy[np.minimum((t * 4800).astype(int), nsym - 1)]
    k = np.hanning(400); k /= k.sum()
    ph = 2 * np.pi * np.cumsum(np.convolve(x, k, mode="same") * 800.0) / 1e6
    iq = 0.5 * np.exp(1j * ph) + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.01
    return iq.astype(np.complex64)[None, :]


def _sig_ssb(O, sg):
    rng = np.random.default_rng(9600)
    T = 1 << 18
    n = np.arange(T)
    x = 0.3 * np.exp(2j * np.pi * 1200.0 * n / 1e6) + 0.2 * np.exp(2j * np.pi * 700.0 * n / 1e6)
    x = x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.005
    return x.astype(np.complex64)[None, :]


# name -> oracle kind, factory args (sps, samp_rate, carrier, filter_width, flag), ports, product factory name + args
RX_CASES = {
    "4fsk_2k_fm": dict(okind=2, args=(5, 1000000, 1700, 3000, 1), nports=3, signal=_sig_4fsk(True),
                       factory="make_gr_demod_4fsk", fargs=(5, 1000000, 1700, 3000, True)),
    "4fsk_2k": dict(okind=2, args=(5, 1000000, 1700, 4000, 0), nports=3, signal=_sig_4fsk(False),
                    factory="make_gr_demod_4fsk", fargs=(5, 1000000, 1700, 4000, False)),
    "qpsk_250k": dict(okind=3, args=(2, 1000000, 1700, 160000, 0), nports=3, signal=_sig_qpsk,
                      factory="make_gr_demod_qpsk", fargs=(2, 1000000, 1700, 160000)),
    "nbfm_2500": dict(okind=1, args=(125, 1000000, 1700, 2500, 0), nports=2, signal=_sig_nbfm,
                      factory="make_gr_demod_nbfm", fargs=(125, 1000000, 1700, 2500)),
    "bpsk_2k": dict(okind=4, args=(250, 1000000, 1700, 2800, 0), nports=4, signal=_sig_digital("bpsk"),
                    factory="make_gr_demod_bpsk", fargs=(250, 1000000, 1700, 2800)),
    "2fsk_2k_fm": dict(okind=5, args=(25, 1000000, 1700, 4000, 1), nports=4, signal=_sig_digital("2fsk"),
                       factory="make_gr_demod_2fsk", fargs=(25, 1000000, 1700, 4000, True)),
    "am_5000": dict(okind=7, args=(125, 1000000, 1700, 5000, 0), nports=2, signal=_sig_am,
                    factory="make_gr_demod_am", fargs=(125, 1000000, 1700, 5000)),
    "gmsk_2k": dict(okind=8, args=(5, 1000000, 1700, 4000, 0), nports=4, signal=_sig_digital("2fsk"),
                    factory="make_gr_demod_gmsk", fargs=(5, 1000000, 1700, 4000)),
    "wbfm_75k": dict(okind=9, args=(125, 1000000, 1700, 75000, 0), nports=2, signal=_sig_wbfm,
                     factory="make_gr_demod_wbfm", fargs=(125, 1000000, 1700, 75000)),
    "m17": dict(okind=10, args=(125, 1000000, 1700, 9000, 0), nports=3, signal=_sig_m17,
                factory="make_gr_demod_m17", fargs=(125, 1000000, 1700, 9000)),
    "ssb_usb": dict(okind=6, args=(125, 1000000, 1700, 2700, 0), nports=2, signal=_sig_ssb,
                    factory="make_gr_demod_ssb", fargs=(125, 1000000, 1700, 2700, 0)),
    "dmr": dict(okind=11, args=(5, 1000000, 0, 0, 0), nports=4, signal=_sig_m17,
                factory="make_gr_demod_dmr", fargs=(5, 1000000)),
}


def _bytes(seed, n):
    return lambda: np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


TX_CASES = {
    "tx_4fsk_2k_fm": dict(okind=101, args=(25, 1000000, 1700, 3500, 1), data=_bytes(9700, 24),
                          factory="make_gr_mod_4fsk", fargs=(25, 1000000, 1700, 3500, True)),
    "tx_qpsk_250k": dict(okind=102, args=(4, 1000000, 1700, 160000, 0), data=_bytes(9701, 600),
                         factory="make_gr_mod_qpsk", fargs=(4, 1000000, 1700, 160000)),
    "tx_bpsk_2k": dict(okind=104, args=(250, 1000000, 1700, 2800, 0), data=_bytes(9702, 12),
                       factory="make_gr_mod_bpsk", fargs=(250, 1000000, 1700, 2800)),
    "tx_2fsk_2k_fm": dict(okind=105, args=(25, 1000000, 1700, 4000, 1), data=_bytes(9703, 12),
                          factory="make_gr_mod_2fsk", fargs=(25, 1000000, 1700, 4000, True)),
    "tx_m17": dict(okind=108, args=(125, 1000000, 1700, 9000, 0), data=_bytes(9704, 60),
                   factory="make_gr_mod_m17", fargs=(125, 1000000, 1700, 9000)),
    "tx_dmr": dict(okind=109, args=(125, 1000000, 1700, 5000, 0), data=_bytes(9705, 120),
                   factory="make_gr_mod_dmr", fargs=(125, 1000000, 1700, 5000)),
}


# ---- blocks outside the Rx / Tx factories (added later in round 2): seeded inputs and the oracle call that defines the answer
def mmdvm_rx_input():
    rng = np.random.default_rng(9800)
    n = 30000
    t = np.arange(n)
    a = 0.4 * np.sin(2 * np.pi * 700 * t / 25000) + 0.2 * np.sin(2 * np.pi * 1900 * t / 25000 + 0.5)
    x = 0.05 * (1 + 0.7 * np.sin(2 * np.pi * 2 * t / 25000)) * np.exp(1j * 2 * np.pi * 2500 * np.cumsum(a) / 25000)
    return (x + 0.001 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)


def mmdvm_tx_input():
    rng = np.random.default_rng(9801)
    n = 12000
    t = np.arange(n)
    return (7000 * np.sin(2 * np.pi * 800 * t / 24000) + rng.integers(-500, 500, n)).astype(np.int16)


def spectrum_input():
    rng = np.random.default_rng(9802)
    n = 4096 + 100
    t = np.arange(n)
    return (0.3 * np.exp(2j * np.pi * 0.2 * t) + 0.02 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)


def dsss_decoder_input():
    rng = np.random.default_rng(9803)
    code = np.array([1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 0, 1])
    chips = np.repeat(np.where(code > 0, 1.0, -1.0), 25)
    bits = rng.integers(0, 2, 12) * 2 - 1
    x = np.concatenate([b * chips for b in bits]) * np.exp(0.7j)
    return (x + 0.4 * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))).astype(np.complex64)


def extra_outputs(O):
    """name -> array, the oracle's answers for the inputs above."""
    import ctypes as C
    out = {}
    o, db, at = O.MmdvmRx(5000).work(mmdvm_rx_input())
    out["mmdvm_rx_int16"] = o
    out["mmdvm_rx_rssi_db"] = db
    out["mmdvm_tx_iq"] = O.MmdvmTx(5000).work(mmdvm_tx_input())
    s = O.Spectrum(4096, O.WIN_BLACKMAN_HARRIS); s.set_enabled(True); s.work(spectrum_input())
    out["spectrum_4096_db"] = s.get()
    x = dsss_decoder_input()
    code = np.array([1, 1, 1, 1, 1, 0, 0, 1, 1, 0, 1, 0, 1], np.int32)
    got = np.zeros(16, np.complex64)
    m = O.lib().qo_dsss_decoder_run(code.ctypes.data_as(C.c_void_p), 13, 25, x.ctypes.data_as(C.c_void_p), len(x), len(x),
                                    got.ctypes.data_as(C.c_void_p), len(got))
    out["dsss_decoder_symbols"] = got[:m].copy()
    return out
