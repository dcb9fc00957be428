#!/bin/bash
# round 2, GPU call c: lean Costas loop, Viterbi v2, NBFM squelch fast path, analog setters, many-channel QPSK geometry
set -u
OUT=gpurun_out/r02_c
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
tail -15 "$OUT/1_gpu_tier.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/2_bench.json" 2> "$OUT/2_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/2_bench.err" | tee -a "$OUT/summary.txt"
