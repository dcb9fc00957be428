#!/bin/bash
# round 2, GPU call t: lean symbol sync with the sample rows prefetched (loads off the loop-carried chain), Costas saturation selects
# in front of the table load: parity tests of every chain that uses either, then cycles / timing
set -u
OUT=gpurun_out/r02_t
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_4fsk.py tests/test_gpu_qpsk.py tests/test_gpu_fast_modes.py tests/test_gpu_m17.py tests/test_gpu_dmr.py tests/test_gpu_bpsk_2fsk.py tests/test_gpu_mixed.py tests/test_golden.py -m gpu -q -x > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
timeout 120 tools/microbench/lone_warp 2>&1 | tail -6 | tee "$OUT/1_lone_warp.txt"
timeout 200 python tools/ss_prof.py cfg2 2>&1 | tee "$OUT/2_ss_prof_cfg2.txt"
cat > /tmp/cfg_time.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench, qradiolink_b200 as q
from qradiolink_b200 import synth
L = q.load_library()
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
C, T = 64, 1 << 22
bases = [synth.burst_4fsk(1000 + i, T) for i in range(4)]
X = synth.batch_on_device(bases, C, seed=4242, device=dev)
blk = q.make_gr_demod_4fsk(5, 1000000, 1700, 3000, True, n_channels=C, max_samples=T)
blk.set_stream(st.cuda_stream); blk.set_overlap(True)
L.qrl_rx_profile(blk._h, 1)
ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), 200, st, torch, warm=5)
blk.join(); blk.sync()
stg = bench.rx_stage_ms(L, blk)
print("cfg2 64ch 2^22 overlapped: ms_per_call=%.4f Msamples/s=%.0f stages=%s" % (ms, C * T / ms / 1e3, {k: round(v[0] / 205, 4) for k, v in stg.items()}), flush=True)
blk.close(); del X
C, T = 256, 1 << 20
X = bench.qpsk_inputs(q, torch, dev, C, T, 2000)
blk = q.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
blk.set_stream(st.cuda_stream)
L.qrl_rx_profile(blk._h, 1)
ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), 3, st, torch, warm=1)
stg = bench.rx_stage_ms(L, blk)
print("QPSK C=%d ms_per_call=%.3f Msamples/s=%.0f stages=%s" % (C, ms, C * T / ms / 1e3, {k: round(v[0] / 4, 3) for k, v in stg.items()}), flush=True)
blk.close()
PY
timeout 300 python /tmp/cfg_time.py 2>&1 | tee "$OUT/3_timing.txt"
