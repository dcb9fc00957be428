#!/usr/bin/env python
"""Small fixed workloads for `ncu --set full` captures (one case per invocation, a couple of calls, no timing):
   python tools/ncu_case.py {cfg2|qpsk|nbfm|tx|pfb|spectrum|mmdvm|dsss|amtx} [channels] [log2 T]
Inputs come from the product's own modulators on the GPU (or torch for the analog case), like bench.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qradiolink_b200 as q  # noqa: E402
from qradiolink_b200 import synth  # noqa: E402


def main():
    case = sys.argv[1]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    calls = 3
    if case == "cfg2":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        T = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 22)
        bases = [synth.burst_4fsk(1000 + i, T) for i in range(4)]
        X = synth.batch_on_device(bases, C, seed=4242, device=dev)
        blk = q.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
        blk.set_stream(stream.cuda_stream)
        if os.environ.get("QRL_CASE_OVERLAP", "1") == "1":
            blk.set_overlap(True)
        for _ in range(calls):
            blk.work_device(X.data_ptr(), T, T)
        blk.join(); blk.sync(); blk.close()
    elif case == "qpsk":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        T = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 18)
        X = bench.qpsk_inputs(q, torch, dev, C, T, 2000)
        blk = q.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
        blk.set_stream(stream.cuda_stream)
        for _ in range(calls):
            blk.work_device(X.data_ptr(), T, T)
        blk.sync(); blk.close()
    elif case == "nbfm":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        T = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 20)
        X = bench.nbfm_inputs(torch, dev, C, T)
        blk = q.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=C, max_samples=T)
        blk.set_stream(stream.cuda_stream)
        for _ in range(calls):
            blk.work_device(X.data_ptr(), T, T)
        blk.sync(); blk.close()
    elif case == "tx":
        C, n = 64, 1024
        tx = q.make_gr_mod_4fsk(25, 1000000, 1700, 3500, True, n_channels=C, max_items=n)
        tx.set_stream(stream.cuda_stream)
        data = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (C, n), dtype=np.uint8)).to(dev)
        for _ in range(calls):
            tx.work_device(data.data_ptr(), n, n)
        tx.sync(); tx.close()
    elif case == "pfb":
        import ctypes as Ct
        M, n_t = 10, 341
        taps = np.zeros(n_t, np.float32)
        L = q.load_library()
        assert L.qrl_firdes_low_pass_2(1.0, 250000.0, 5000.0, 2000.0, 60.0, 5, taps.ctypes.data_as(Ct.c_void_p), n_t) == n_t
        N = 1 << 25
        x = torch.view_as_complex(torch.randn((N, 2), device=dev) * 0.3)
        ch = q.PfbChannelizer(M, taps, max_in=N); ch.set_stream(stream.cuda_stream)
        for _ in range(calls):
            ch.work_device(x.data_ptr(), N)
        ch.sync(); ch.close()
        z = torch.view_as_complex(torch.randn((M, N // M, 2), device=dev) * 0.3)
        sy = q.PfbSynthesizer(M, taps * 10.0, max_in=N // M); sy.set_stream(stream.cuda_stream)
        for _ in range(calls):
            sy.work_device(z.data_ptr(), N // M, N // M)
        sy.sync(); sy.close()
    elif case == "spectrum":
        S, N = 64, 32768
        x = torch.view_as_complex(torch.randn((S, 2 * N, 2), device=dev) * 0.2)
        sp = q.Spectrum(N, 5, n_streams=S, max_samples=2 * N)
        sp.set_stream(stream.cuda_stream); sp.set_enabled(True)
        for _ in range(calls):
            sp.work_device(x.data_ptr(), N + 1, x.shape[1])
            assert sp.get_fft_data() is not None
        sp.close()
    elif case == "mmdvm":
        import ctypes as Ct
        N = 1 << 24
        x = torch.view_as_complex(torch.randn((N, 2), device=dev) * 0.05)
        dem = q.MmdvmDemod(7, 5000, max_in=N)
        dem.channelizer.set_stream(stream.cuda_stream); dem.channels.set_stream(stream.cuda_stream)
        L = q.load_library()
        cnt = Ct.c_long()
        for _ in range(calls):
            pf = dem.channelizer
            assert L.qrl_pfb_work(pf._h, Ct.c_void_p(x.data_ptr()), N, 0, 1, Ct.byref(cnt)) == 0
            ptr, stride, items = pf.out_device()
            dem.channels.work_device(ptr, items, stride)
        torch.cuda.synchronize(); dem.close()
    elif case == "dsss":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        T = 1 << 22
        X = bench.nbfm_inputs(torch, dev, C, T)            # any band-limited signal: the chain's cost does not depend on lock
        blk = q.make_gr_demod_dsss(n_channels=C, max_samples=T)
        blk.set_stream(stream.cuda_stream)
        for _ in range(calls):
            blk.work_device(X.data_ptr(), T, T)
        blk.sync(); blk.close()
    elif case == "amtx":
        C, n = 16, 4000
        tx = q.make_gr_mod_am(125, 1000000, 1700, 5000, n_channels=C, max_items=n)
        tx.set_stream(stream.cuda_stream)
        au = (torch.randn((C, n), device=dev) * 0.3).contiguous()
        for _ in range(calls):
            tx.work_device(au.data_ptr(), n, n)
        tx.sync(); tx.close()
    torch.cuda.synchronize()
    print("case %s done" % case)


if __name__ == "__main__":
    main()
