#!/usr/bin/env python
"""Cycle breakdown of the symbol-sync loop warp (profiling build, -DQRL_SS_PROF; never the shipped library):
   make -C tools/microbench prof   (here: nvcc cross-compiles)
   python tools/ss_prof.py {cfg2|qpsk} [channels] [log2 T]       (on the GPU)
Loads tools/microbench/libqrl_b200_prof.so in place of the product library, runs a few calls, prints where the loop warp of
CTA 0 spent its cycles per window."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qradiolink_b200.lib as qlib  # noqa: E402
qlib._LIB_PATH = os.path.join(ROOT, "tools", "microbench", "libqrl_b200_prof.so")
import bench  # noqa: E402
import qradiolink_b200 as q  # noqa: E402
from qradiolink_b200 import synth  # noqa: E402


def main():
    case = sys.argv[1]
    dev = torch.device("cuda", 0)
    L = q.load_library()
    L.qrl_debug_ss_prof.argtypes = [ctypes.c_void_p]
    out = (ctypes.c_longlong * 16)()
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    if case == "cfg2":
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
        T = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 22)
        bases = [synth.burst_4fsk(1000 + i, T) for i in range(4)]
        X = synth.batch_on_device(bases, C, seed=4242, device=dev)
        blk = q.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
    else:
        C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        T = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 20)
        X = bench.qpsk_inputs(q, torch, dev, C, T, 2000)
        blk = q.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
    blk.set_stream(stream.cuda_stream)
    if case == "cfg2":
        blk.set_overlap(True)
    for it in range(4):
        blk.work_device(X.data_ptr(), T, T)
        if case == "cfg2":
            blk.join()
        blk.sync()
        assert L.qrl_debug_ss_prof(ctypes.byref(out)) == 0
        v = list(out)
        w = max(v[5], 1)
        tot = sum(v[0:5])
        print("%s call %d: windows %d, symbols %d (%.1f per window), uniform rounds %.2f per window; cycles per window: "
              "wait hand-off %.0f, wait window %.0f, uniform rounds %.0f, stragglers %.0f, hand-off %.0f, total %.0f (%.1f per symbol)"
              % (case, it, v[5], v[6], v[6] / w, v[7] / w, v[0] / w, v[1] / w, v[2] / w, v[3] / w, v[4] / w, tot / w, tot / max(v[6], 1)))
        nl = max(v[11], 1)
        print("    launches %d; per launch (globaltimer, us): entry -> loop warp starts %.2f, entry -> first window there %.2f, entry -> loop warp done %.2f"
              % (v[11], v[8] / nl / 1e3, v[9] / nl / 1e3, v[10] / nl / 1e3))
    blk.close()


if __name__ == "__main__":
    main()
