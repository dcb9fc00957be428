"""config 5 (64 ch 4FSK-2k-FM TX) timing helper: python tools/tx_case.py [k]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qradiolink_b200 as q  # noqa: E402
import tools.bench_configs as b  # noqa: E402

dev = torch.device("cuda", 0); rng = np.random.default_rng(7)
C, n = 64, 1024
tx = q.make_gr_mod_4fsk(25, 1000000, 1700, 3500, True, n_channels=C, max_items=n)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); tx.set_stream(stream.cuda_stream)
data = torch.from_numpy(rng.integers(0, 256, (C, n), dtype=np.uint8)).to(dev)
fn = lambda: tx.work_device(data.data_ptr(), n, n)  # noqa: E731
ms = b.timed(fn, int(sys.argv[1]) if len(sys.argv) > 1 else 5, stream); nout = n * 4000
print(json.dumps({"config": "64ch 4FSK TX (config 5)", "ms_per_call": round(ms, 4), "Msamples_per_s": round(C * nout / ms / 1e3, 1),
                  "frac_of_hbm_peak": round(C * nout * 8 / ms / 1e6 / b.HBM, 3)}))
