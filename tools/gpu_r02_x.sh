#!/bin/bash
# round 2, GPU call x: MMDVM zero-idle tests, compute-sanitizer over the kernels changed late in the round (agc/Costas hand-off ring,
# prefetched symbol-sync rows, table selects, MMDVM zero ranges), then the bench line as the driver runs it
set -u
OUT=gpurun_out/r02_x
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_mmdvm.py tests/test_gpu_dmr_tx.py -m gpu -q -x > "$OUT/0_tests.log" 2>&1
echo "mmdvm / dmr tx tests exit $?" | tee "$OUT/summary.txt"
tail -2 "$OUT/0_tests.log" | tee -a "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_qpsk.py::test_qpsk_chunked_stream" "tests/test_gpu_bpsk_2fsk.py::test_bpsk_2k_parity" "tests/test_gpu_4fsk.py::test_4fsk_fm_chunk_invariance_and_ragged" "tests/test_gpu_4fsk.py::test_overlapped_calls_with_changing_length" "tests/test_gpu_m17.py" "tests/test_gpu_mmdvm.py::test_tx_zero_idle_bursts" -m gpu -q -x > "$OUT/1_memcheck.log" 2>&1
echo "memcheck exit $?" | tee -a "$OUT/summary.txt"
grep -E "ERROR SUMMARY|passed|failed" "$OUT/1_memcheck.log" | tail -3 | tee -a "$OUT/summary.txt"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_qpsk.py::test_qpsk_parity_and_frames" "tests/test_gpu_bpsk_2fsk.py::test_bpsk_2k_parity" "tests/test_gpu_4fsk.py::test_4fsk_fm_parity_single_call" -m gpu -q -x > "$OUT/2_racecheck.log" 2>&1
echo "racecheck exit $?" | tee -a "$OUT/summary.txt"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" "$OUT/2_racecheck.log" | tail -4 | tee -a "$OUT/summary.txt"
SECONDS=0
timeout 1200 python bench.py --steps 20 --warmup 5 > "$OUT/3_bench.json" 2> "$OUT/3_bench.err"
echo "bench exit $? after ${SECONDS}s" | tee -a "$OUT/summary.txt"
tail -2 "$OUT/3_bench.err" | tee -a "$OUT/summary.txt"
python - <<'PY' | tee -a "$OUT/summary.txt"
import json
d = json.loads(open("gpurun_out/r02_x/3_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms_per_step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "whole", round(d["roofline"]["whole_chain_frac"], 3))
for k, v in d["configs"].items():
    if isinstance(v, dict) and "value" in v: print(k, round(v["value"]), v.get("unit"), "ms", round(v.get("ms_per_call", 0), 3), "parity", (v.get("parity_vs_oracle") or {}).get("bits_equal", (v.get("parity_vs_oracle") or {}).get("float_rms_max")))
PY
