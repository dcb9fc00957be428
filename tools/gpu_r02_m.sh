#!/bin/bash
# round 2, GPU call m: PFB channelizer after the bank-conflict fix (tests, duration, full capture)
set -u
OUT=gpurun_out/r02_m
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_pfb.py tests/test_gpu_mmdvm.py -m gpu -q > "$OUT/0_tests.log" 2>&1
echo "tests exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file "$OUT/launches_pfb.csv" python tools/ncu_case.py pfb > "$OUT/ncu_pfb.log" 2>&1
echo "ncu pfb exit $?" | tee -a "$OUT/summary.txt"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:pfb_chan_kernel --launch-skip 1 -c 1 -o "$OUT/ncu_pfb_chan" python tools/ncu_case.py pfb > "$OUT/ncu_pfb_full.log" 2>&1
echo "ncu full exit $?" | tee -a "$OUT/summary.txt"
python tools/ncu_summary.py "$OUT/ncu_summary.csv" "$OUT"/ncu_pfb_chan.ncu-rep | tail -1
