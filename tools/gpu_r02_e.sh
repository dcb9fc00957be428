#!/bin/bash
# round 2, GPU call e: branch-free symbol-sync / Costas loops, tiled low-rate filters, device framing additions
set -u
OUT=gpurun_out/r02_e
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
tail -12 "$OUT/1_gpu_tier.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/2_bench.json" 2> "$OUT/2_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/2_bench.err" | tee -a "$OUT/summary.txt"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 re=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:"$re" --launch-skip "$skip" -c 1 -o "$OUT/ncu_$name" python tools/ncu_case.py "$@" > "$OUT/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$OUT/summary.txt"; }
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap symsync_4fsk "symsync_kernel"     1 cfg2 64 22
cap agc_costas  "agc_costas_kernel"   1 qpsk 256 18
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap chanfilt "fir_ccf_ring_tiled_kernel"  1 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap qdemod "qdemod_fir_fff_tiled_kernel"  1 cfg2 64 22
