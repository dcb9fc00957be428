#!/bin/bash
# round 2, GPU call w: whole GPU tier after the late loop changes (3-deep agc/Costas ring, prefetched symbol-sync rows, Costas table
# selects, QPSK without a partition, 2 slices per overlapped call) + where a symbol-sync launch spends its start-up
set -u
OUT=gpurun_out/r02_w
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
tail -4 "$OUT/1_gpu_tier.log" | tee -a "$OUT/summary.txt"
timeout 200 python tools/ss_prof.py cfg2 2>&1 | grep -v Warning | tee "$OUT/2_ss_prof_cfg2.txt"
