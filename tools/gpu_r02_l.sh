#!/bin/bash
# round 2, GPU call l: launch lists (durations) of the new chains: MMDVM demodulator, DSSS receive chain, AM modulator; AM TX test
set -u
OUT=gpurun_out/r02_l
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tx.py -m gpu -q -k "tx_am" > "$OUT/0_am_test.log" 2>&1
echo "am tx test exit $?" | tee "$OUT/summary.txt"
tail -5 "$OUT/0_am_test.log" | tee -a "$OUT/summary.txt"
for c in mmdvm dsss amtx; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/launches_$c.csv" python tools/ncu_case.py $c > "$OUT/ncu_$c.log" 2>&1
  echo "ncu $c exit $?" | tee -a "$OUT/summary.txt"
done
