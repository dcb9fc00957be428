#!/bin/bash
# round 2, GPU call d: SSB gain fix, device gr_deframer_bb, analog partition; ncu of the rewritten loop kernels
set -u
OUT=gpurun_out/r02_d
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_setters.py tests/test_gpu_framing.py tests/test_gpu_ssb.py tests/test_gpu_nbfm.py tests/test_gpu_am.py tests/test_gpu_wbfm.py -m gpu -q > "$OUT/1_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -5 "$OUT/1_tests.log" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/2_bench.json" 2> "$OUT/2_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { local name=$1 re=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:"$re" --launch-skip "$skip" -c 1 -o "$OUT/ncu_$name" python tools/ncu_case.py "$@" > "$OUT/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$OUT/summary.txt"; }
cap agc_costas  "agc_costas_kernel"   1 qpsk 256 18
cap symsync_q   "symsync_kernel"      1 qpsk 256 18
cap viterbi_q   "viterbi_k7_kernel"   1 qpsk 256 18
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap viterbi_4fsk "viterbi_k7_kernel"  1 cfg2 64 22
QRL_NSUB=1 QRL_CASE_OVERLAP=0 cap symsync_4fsk "symsync_kernel"     1 cfg2 64 22
cap nbfm_audio  "nbfm_audio_kernel"   1 nbfm 64 20
