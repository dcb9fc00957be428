#!/bin/bash
# round 2, GPU call u: the Viterbi decoder on a third SM partition of its own (A/B against sharing the loop partition)
set -u
OUT=gpurun_out/r02_u
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_4fsk.py tests/test_gpu_qpsk.py tests/test_gpu_mixed.py -m gpu -q -x > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
(timeout 300 python tools/cfg_time.py; QRL_FEC_SMS=0 timeout 300 python tools/cfg_time.py; QRL_FEC_SMS=16 timeout 300 python tools/cfg_time.py; QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py cfg2) 2>&1 | grep -v Warning | tee "$OUT/1_ab.txt"
