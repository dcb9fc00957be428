#!/bin/bash
# round 2, GPU call s: agc2 / Costas stage with a 3-deep hand-off ring (QPSK / BPSK parity, QPSK timing per stage) and the
# cycle breakdown of the symbol-sync loop warp (profiling build)
set -u
OUT=gpurun_out/r02_s
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_qpsk.py tests/test_gpu_bpsk_2fsk.py tests/test_gpu_mixed.py -m gpu -q -x > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
cat > /tmp/qpsk_time.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench, qradiolink_b200 as q
L = q.load_library()
dev = torch.device('cuda', 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for C, lt in ((256, 20), (1024, 18)):
    T = 1 << lt
    X = bench.qpsk_inputs(q, torch, dev, C, T, 2000)
    blk = q.make_gr_demod_qpsk(2, 1000000, 1700, 160000, n_channels=C, max_samples=T)
    blk.set_stream(st.cuda_stream)
    L.qrl_rx_profile(blk._h, 1)
    ms = bench.timed_calls(lambda: blk.work_device(X.data_ptr(), T, T), 3, st, torch, warm=1)
    stg = bench.rx_stage_ms(L, blk)
    print("QPSK C=%d T=2^%d ms_per_call=%.3f Msamples/s=%.0f stages=%s" % (C, lt, ms, C * T / ms / 1e3, {k: round(v[0] / 4, 3) for k, v in stg.items()}), flush=True)
    blk.close(); del X
PY
timeout 300 python /tmp/qpsk_time.py 2>&1 | tee "$OUT/1_qpsk_time.txt"
timeout 200 python tools/ss_prof.py cfg2 2>&1 | tee "$OUT/2_ss_prof_cfg2.txt"
timeout 200 python tools/ss_prof.py qpsk 256 19 2>&1 | tee "$OUT/3_ss_prof_qpsk.txt"
