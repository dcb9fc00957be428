#!/bin/bash
# round 2, closing GPU call (after the last session's loop / filter changes): whole GPU tier, the bench line as the driver runs it (+ reference arm), the launch list of the same command,
# ncu --set full captures of the kernels the roofline discussion names
set -u
OUT=gpurun_out/r02_final2
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
tail -6 "$OUT/1_gpu_tier.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/2_bench_reference_arm.json" 2> "$OUT/2_bench_reference_arm.err"
echo "reference arm exit $?" | tee -a "$OUT/summary.txt"
timeout 1200 python bench.py --steps 20 --warmup 5 > "$OUT/3_bench.json" 2> "$OUT/3_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/3_bench.err" | tee -a "$OUT/summary.txt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_|symsync|viterbi|qdemod|hist_|frontend" -c 3000 --csv --log-file "$OUT/4_launches_step.csv" \
    python bench.py --steps 2 --warmup 3 --headline-only --no-cpu --no-parity > "$OUT/4_launches_step.log" 2>&1
echo "launch list exit $?" | tee -a "$OUT/summary.txt"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 re=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:"$re" --launch-skip "$skip" -c 1 -o "$OUT/ncu_$name" python tools/ncu_case.py "$@" > "$OUT/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$OUT/summary.txt"; }
cap fir_poly   "fir_decim_poly_kernel"      2 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap symsync_4fsk "symsync_kernel"   1 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap viterbi      "viterbi_k7_kernel" 1 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap chan_filter  "fir_ccf_ring_tiled_kernel" 1 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap demod_rrc    "qdemod_fir_fff_tiled_kernel" 1 cfg2 64 22
QRL_NSUB=1 cap agc_costas "agc_costas_kernel"          1 qpsk 256 18
QRL_NSUB=1 cap symsync_qpsk "symsync_kernel"           1 qpsk 256 18
cap am_fir     "fir_ccf_ring_tiled_kernel"  1 amtx
python tools/ncu_summary.py "$OUT/5_ncu_full_summary.csv" "$OUT"/ncu_*.ncu-rep 2>&1 | tail -1 | tee -a "$OUT/summary.txt"
ls -la "$OUT" | tail -30 >> "$OUT/summary.txt"
