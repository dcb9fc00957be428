"""GPU tier: run-time setters of the analog demodulators (set_squelch, set_filter_width, set_ctcss, set_agc_attack /
set_agc_decay, set_gain -- gr_demod_nbfm.cpp:82-121, gr_demod_ssb.cpp:89-121, gr_demod_am.cpp:84-107, gr_demod_wbfm.cpp:77-91)
through the C ABI (qrl_rx_set_param), mid-stream, against the CPU oracle's restatement of the same setters: all ports bit-identical
before and after the change."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fm_signal(rng, C, T, dev_hz):
    n = np.arange(T)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        audio = 0.6 * np.sin(2 * np.pi * (900.0 + 70 * c) * n / 1e6) + 0.3 * np.sin(2 * np.pi * 2100.0 * n / 1e6)
        ph = 2 * np.pi * dev_hz * np.cumsum(audio) / 1e6
        amp = 0.5 * (1.0 + 0.9 * np.sin(2 * np.pi * 1.7 * n / 1e6 + c))          # deep fades: a raised squelch opens and closes
        x = amp * np.exp(1j * ph) + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.003
        X[c] = x.astype(np.complex64)
    return X


def am_signal(rng, C, T):
    n = np.arange(T)
    X = np.zeros((C, T), np.complex64)
    for c in range(C):
        aud = 0.5 * np.sin(2 * np.pi * (700.0 + 130.0 * c) * n / 1e6) + 0.3 * np.sin(2 * np.pi * 1900.0 * n / 1e6 + 0.4)
        x = 0.4 * (1.0 + 0.8 * aud) * np.exp(2j * np.pi * (rng.uniform(-300, 300) * n / 1e6 + rng.uniform(0, 1)))
        x = x * (1.0 + 0.9 * np.sin(2 * np.pi * 2.3 * n / 1e6)) * 0.6
        X[c] = (x + (rng.standard_normal(T) + 1j * rng.standard_normal(T)) * 0.004).astype(np.complex64)
    return X


def run_case(qrl, oracle, make, okind, args, oflag, X, steps):
    """steps: list of (n_samples, [(PARAM key, value), ...]) -- the setters are applied BEFORE the chunk on both sides."""
    C, T = X.shape
    blk = make(*args, n_channels=C, max_samples=max(n for n, _ in steps))
    rxs = [oracle.Rx(okind, *args[:4], oflag) for _ in range(C)]
    lo = 0
    for n, sets in steps:
        for key, value in sets:
            blk.set_param(key, value)
            for rx in rxs:
                rx.set_param(key, value)
        blk.work(X[:, lo:lo + n])
        got = [blk.read_port(p) for p in range(2)]
        for c in range(C):
            rxs[c].work(X[c, lo:lo + n])
            for p in range(2):
                want = rxs[c].port(p)
                assert len(got[p][c]) == len(want), (lo, c, p, len(got[p][c]), len(want))
                assert np.array_equal(got[p][c], want), (lo, c, p)
        lo += n
    assert lo <= T
    blk.close()


def test_nbfm_setters(qrl, oracle):
    P = qrl.PARAM
    X = fm_signal(np.random.default_rng(71), 3, 900000, 1800.0)
    run_case(qrl, oracle, qrl.make_gr_demod_nbfm, oracle.DEMOD_NBFM, (125, 1000000, 1700, 2500), 0, X,
             [(200000, []), (200000, [(P.SQUELCH_DB, -9.0)]), (150000, [(P.FILTER_WIDTH, 5000)]), (100001, [(P.CTCSS, 0.0)]),
              (200000, [(P.SQUELCH_DB, -140.0), (P.FILTER_WIDTH, 3000)])])


@pytest.mark.parametrize("sb", [0, 1])
def test_ssb_setters(qrl, oracle, sb):
    P = qrl.PARAM
    T = 900000
    n = np.arange(T)
    rng = np.random.default_rng(72 + sb)
    X = np.zeros((2, T), np.complex64)
    for c in range(2):
        sign = -1.0 if sb else 1.0
        x = 0.05 * np.exp(2j * np.pi * sign * (700 + 150 * c) * n / 1e6) * (1 + 0.9 * np.sin(2 * np.pi * 3 * n / 1e6))
        x = x + 0.03 * np.exp(2j * np.pi * sign * 1900 * n / 1e6) + 0.002 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))
        X[c] = x.astype(np.complex64)
    run_case(qrl, oracle, qrl.make_gr_demod_ssb, oracle.DEMOD_SSB, (125, 1000000, 1700, 2700, sb), sb, X,
             [(250000, []), (250000, [(P.AGC_ATTACK, 0.02), (P.AGC_DECAY, 0.3), (P.GAIN, 0.5)]), (200000, [(P.FILTER_WIDTH, 2200)]),
              (200000, [(P.SQUELCH_DB, -27.0)])])


def test_am_setters(qrl, oracle):
    P = qrl.PARAM
    X = am_signal(np.random.default_rng(73), 2, 800000)
    run_case(qrl, oracle, qrl.make_gr_demod_am, oracle.DEMOD_AM, (125, 1000000, 1700, 5000), 0, X,
             [(200000, []), (200000, [(P.FILTER_WIDTH, 3500)]), (200000, [(P.AGC_ATTACK, 0.5), (P.AGC_DECAY, 0.01)]), (200000, [(P.SQUELCH_DB, -6.0)])])


def test_wbfm_setters(qrl, oracle):
    P = qrl.PARAM
    X = fm_signal(np.random.default_rng(74), 2, 600000, 50000.0)
    run_case(qrl, oracle, qrl.make_gr_demod_wbfm, oracle.DEMOD_WBFM, (125, 1000000, 1700, 75000), 0, X,
             [(200000, []), (200000, [(P.FILTER_WIDTH, 60000)]), (200000, [(P.SQUELCH_DB, -8.0)])])


def test_setters_refused_where_the_reference_has_none(qrl):
    P = qrl.PARAM
    blk = qrl.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=1, max_samples=65536)
    for key in (P.SQUELCH_DB, P.FILTER_WIDTH, P.CTCSS, P.AGC_ATTACK, P.GAIN):
        with pytest.raises(qrl.QrlError):
            blk.set_param(key, 1.0)
    nb = qrl.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=1, max_samples=65536)
    with pytest.raises(qrl.QrlError):
        nb.set_param(P.GAIN, 0.5)            # gr_demod_nbfm has no set_gain
    am = qrl.make_gr_demod_am(125, 1000000, 1700, 5000, n_channels=1, max_samples=65536)
    with pytest.raises(qrl.QrlError):
        am.set_param(P.CTCSS, 88.5)          # only gr_demod_nbfm has a tone squelch


def test_ssb_agc_with_hot_input(qrl, oracle):
    """Input levels above the AGC reference drive the gain below it: the one region where agc2_cc's signed rate compare and
    agc2_ff's fabsf compare differ (oracle/qrl_oracle.c agc2_step), with attack != decay and deep fades."""
    P = qrl.PARAM
    T = 700000
    n = np.arange(T)
    rng = np.random.default_rng(79)
    X = np.zeros((2, T), np.complex64)
    for c in range(2):
        fade = np.where((n // 90000) % 2 == 0, 1.0, 0.01)
        x = (3.0 + c) * fade * np.exp(2j * np.pi * (800 + 150 * c) * n / 1e6) * (1 + 0.5 * np.sin(2 * np.pi * 5 * n / 1e6))
        X[c] = (x + 0.002 * (rng.standard_normal(T) + 1j * rng.standard_normal(T))).astype(np.complex64)
    run_case(qrl, oracle, qrl.make_gr_demod_ssb, oracle.DEMOD_SSB, (125, 1000000, 1700, 2700, 0), 0, X,
             [(100000, []), (300000, [(P.AGC_ATTACK, 0.5), (P.AGC_DECAY, 0.01)]), (300000, [(P.AGC_ATTACK, 0.01), (P.AGC_DECAY, 0.6)])])


def test_nbfm_tone_squelch(qrl, oracle):
    """gr_demod_nbfm::set_ctcss(f) (gr_demod_nbfm.cpp:97-121): analog::ctcss_squelch_ff in front of a band-pass audio filter, gating: the
    audio port only carries items while the 88.5 Hz tone is present (decision once per 8000-item block, 160-item ramps); then a change of
    tone mid-stream (new Goertzel filters, state kept), then set_ctcss(0).  Channel 0 carries the tone, channel 1 a wrong one, channel 2
    none.  Ports identical to the oracle's restatement for every call."""
    P = qrl.PARAM
    n_aud = 8000 * 5
    t = np.arange(n_aud)
    iqs = []
    for c, tone in enumerate((88.5, 100.0, 0.0)):
        m = oracle.Tx(oracle.MOD_NBFM, 20, 1000000, 1700, 2500, 0)
        if tone:
            m.set_param(7, tone)
        au = (0.4 * np.sin(2 * np.pi * (800 + 150 * c) * t / 8000)).astype(np.float32)
        iqs.append(m.work(au))
    X = np.stack(iqs)
    X = (X * 0.5 + 0.002 * (np.random.default_rng(75).standard_normal(X.shape) + 1j * np.random.default_rng(76).standard_normal(X.shape))).astype(np.complex64)
    steps = [(1000000, [(P.CTCSS, 88.5)]), (1500000, []), (700001, [(P.CTCSS, 100.0)]), (1299999, []), (500000, [(P.CTCSS, 0.0)])]
    assert sum(n for n, _ in steps) == X.shape[1]
    C = X.shape[0]
    blk = qrl.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=C, max_samples=1500000)
    rxs = [oracle.Rx(oracle.DEMOD_NBFM, 125, 1000000, 1700, 2500, 0) for _ in range(C)]
    lo, total = 0, np.zeros(C, int)
    for n, sets in steps:
        for key, value in sets:
            blk.set_param(key, value)
            for rx in rxs:
                rx.set_param(key, value)
        blk.work(X[:, lo:lo + n])
        got = [blk.read_port(p) for p in range(2)]
        for c in range(C):
            rxs[c].work(X[c, lo:lo + n])
            for p in range(2):
                want = rxs[c].port(p)
                assert len(got[p][c]) == len(want), (lo, c, p, len(got[p][c]), len(want))
                assert np.array_equal(got[p][c], want), (lo, c, p)
            total[c] += len(got[1][c])
        lo += n
    # 88.5 Hz on channel 0 opens during the first 2.5 s; the 100 Hz phase opens channel 1 instead; the last call passes everything
    assert total[0] > 8000 and total[1] > 8000 and total[2] == 500000 * 8 // 1000, total
