#!/bin/bash
# round 2, last GPU call: lean complex symbol sync + split epilogue (QPSK): parity tests, then timing against the generic loop
set -u
OUT=gpurun_out/r02_zz
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 150 python -m pytest tests/test_gpu_qpsk.py tests/test_gpu_mixed.py "tests/test_golden.py" -m gpu -q -x -k "qpsk or mixed" > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -2 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
(timeout 60 python tools/cfg_time.py qpsk; QRL_QPSK_LEAN_SS=0 timeout 60 python tools/cfg_time.py qpsk) 2>&1 | grep -v Warning | tee "$OUT/1_ab.txt"
