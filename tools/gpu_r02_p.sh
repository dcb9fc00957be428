#!/bin/bash
# round 2, GPU call p: NBFM with the de-emphasis recurrence split off (tests, timing with and without the split)
set -u
OUT=gpurun_out/r02_p
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_nbfm.py tests/test_gpu_setters.py tests/test_gpu_mixed.py tests/test_golden.py -m gpu -q > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
cat > /tmp/nbfm_time.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench, qradiolink_b200 as q
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for C in (1, 64, 256):
    T = 1 << 22
    X = bench.nbfm_inputs(torch, dev, C, T)
    blk = q.make_gr_demod_nbfm(125, 1000000, 1700, 2500, n_channels=C, max_samples=T)
    blk.set_stream(st.cuda_stream)
    ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), 5, st, torch, warm=2)
    print("C=%d ms_per_call=%.3f Msamples/s=%.0f" % (C, ms, C * T / ms / 1e3), flush=True)
    blk.close(); del X
PY
timeout 600 python /tmp/nbfm_time.py 2>&1 | tee "$OUT/1_split.txt"
QRL_NBFM_NO_SPLIT=1 timeout 600 python /tmp/nbfm_time.py 2>&1 | tee "$OUT/2_nosplit.txt"
