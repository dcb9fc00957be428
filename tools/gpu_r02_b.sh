#!/bin/bash
# round 2, GPU call b: whole GPU tier, the new bench line (all BASELINE configs), ncu --set full captures of the kernels that had none
set -u
OUT=gpurun_out/r02_b
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/1_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee "$OUT/summary.txt"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/2_bench.json" 2> "$OUT/2_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
head -c 1500 "$OUT/2_bench.json" | tee -a "$OUT/summary.txt"; echo | tee -a "$OUT/summary.txt"
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launch-skip, case args...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:"$re" --launch-skip "$skip" -c 1 -o "$OUT/ncu_$name" python tools/ncu_case.py "$@" > "$OUT/ncu_$name.log" 2>&1
  echo "ncu $name exit $?" | tee -a "$OUT/summary.txt"
}
cap fir_poly    "fir_decim_poly_kernel"     2 cfg2 64 22
cap agc_costas  "agc_costas_kernel"         1 qpsk 256 18
cap symsync_q   "symsync_kernel"            1 qpsk 256 18
cap viterbi_q   "viterbi_k7_kernel"         1 qpsk 256 18
cap fir_d2      "fir_decim2_kernel"         1 qpsk 256 18
cap nbfm_audio  "nbfm_audio_kernel"         1 nbfm 64 20
cap interp      "interp_fir_ccf_rt_kernel"  9 tx
cap pfb_chan    "pfb_chan_kernel"           1 pfb
cap pfb_synth   "pfb_synth_kernel"          1 pfb
ls -la "$OUT" | tee -a "$OUT/summary.txt"
