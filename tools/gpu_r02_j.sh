#!/bin/bash
# round 2, GPU call j: DSSS receive chain + modulator (first run on hardware), under compute-sanitizer first
set -u
OUT=gpurun_out/r02_j
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dsss.py -m gpu -q -x -k "ragged or tx_dsss" > "$OUT/0_memcheck.log" 2>&1
echo "memcheck exit $?" | tee "$OUT/summary.txt"
tail -6 "$OUT/0_memcheck.log" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_dsss.py -m gpu -q > "$OUT/1_dsss.log" 2>&1
echo "dsss tests exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/1_dsss.log" | tee -a "$OUT/summary.txt"
