#!/usr/bin/env python
"""Quick device-resident timing of the two latency-bound configurations (A/B runs of environment knobs; not the bench):
   python tools/cfg_time.py [cfg2[:channels]] [qpsk[:channels]]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import qradiolink_b200 as q  # noqa: E402
from qradiolink_b200 import synth  # noqa: E402


def main():
    args = sys.argv[1:] or ["cfg2", "qpsk"]
    which = [a.split(":")[0] for a in args]
    chans = {a.split(":")[0]: int(a.split(":")[1]) for a in args if ":" in a}
    L = q.load_library()
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("QRL_"))
    if "cfg2" in which:
        C, T = chans.get("cfg2", 64), 1 << 22
        bases = [synth.burst_4fsk(1000 + i, T) for i in range(4)]
        X = synth.batch_on_device(bases, C, seed=4242, device=dev)
        blk = q.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
        blk.set_stream(st.cuda_stream); blk.set_overlap(True)
        L.qrl_rx_profile(blk._h, 1)
        k = max(20, 200 * 64 // C)
        ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), k, st, torch, warm=5)
        blk.join(); blk.sync()
        stg = bench.rx_stage_ms(L, blk)
        print("[%s] cfg2 %dch 2^22 overlapped: ms_per_call=%.4f Msamples/s=%.0f stages=%s"
              % (tag, C, ms, C * T / ms / 1e3, {kk: round(v[0] / (k + 5), 4) for kk, v in stg.items()}), flush=True)
        blk.close(); del X
    if "qpsk" in which:
        C, T = chans.get("qpsk", 256), 1 << 20
        X = bench.qpsk_inputs(q, torch, dev, C, T, 2000)
        blk = q.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
        blk.set_stream(st.cuda_stream)
        L.qrl_rx_profile(blk._h, 1)
        ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), 3, st, torch, warm=1)
        stg = bench.rx_stage_ms(L, blk)
        print("[%s] QPSK C=%d 2^20: ms_per_call=%.3f Msamples/s=%.0f stages=%s"
              % (tag, C, ms, C * T / ms / 1e3, {k: round(v[0] / 4, 3) for k, v in stg.items()}), flush=True)
        blk.close()


if __name__ == "__main__":
    main()
