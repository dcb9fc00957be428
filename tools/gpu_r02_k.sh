#!/bin/bash
# round 2, GPU call k: DSSS, MMDVM front end (first hardware run of qrl_mmdvm_*), spectrum; one-GPU tier + bench
set -u
OUT=gpurun_out/r02_k
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
tail -12 "$OUT/1_gpu_tier.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/2_bench.json" 2> "$OUT/2_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/2_bench.err" | tee -a "$OUT/summary.txt"
