#!/bin/bash
# round 2, GPU call n: TX interpolator with several tiles per CTA (tests, duration, full capture)
set -u
OUT=gpurun_out/r02_n
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tx.py tests/test_golden.py -m gpu -q > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file "$OUT/launches_tx.csv" python tools/ncu_case.py tx > "$OUT/ncu_tx.log" 2>&1
echo "ncu tx exit $?" | tee -a "$OUT/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:interp_fir_ccf_rt_kernel --launch-skip 9 -c 1 -o "$OUT/ncu_interp" python tools/ncu_case.py tx > "$OUT/ncu_interp_full.log" 2>&1
echo "ncu full exit $?" | tee -a "$OUT/summary.txt"
python tools/ncu_summary.py "$OUT/ncu_summary.csv" "$OUT"/ncu_interp.ncu-rep | tail -1
