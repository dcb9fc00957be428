#!/bin/bash
# round 2, GPU call q: compute-sanitizer over the kernels added late in the round (memcheck on their tests, racecheck on the shared-memory ones)
set -u
OUT=gpurun_out/r02_q
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mmdvm.py tests/test_gpu_spectrum.py tests/test_gpu_nbfm.py tests/test_gpu_dmr_tx.py "tests/test_gpu_setters.py::test_nbfm_setters" "tests/test_gpu_tx.py::test_tx_am_modulator" "tests/test_gpu_tx.py::test_tx_nbfm_set_ctcss" "tests/test_gpu_4fsk.py::test_sc16_ingest_equals_the_float_path" "tests/test_gpu_4fsk.py::test_sc8_ingest_equals_the_float_path" -m gpu -q -x > "$OUT/0_memcheck.log" 2>&1
echo "memcheck exit $?" | tee "$OUT/summary.txt"
grep -E "ERROR SUMMARY|passed|failed" "$OUT/0_memcheck.log" | tail -4 | tee -a "$OUT/summary.txt"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_spectrum.py::test_spectrum_set_fft_size_and_device_input" "tests/test_gpu_mmdvm.py::test_whole_demodulator_behind_the_channelizer" "tests/test_gpu_nbfm.py" "tests/test_gpu_wbfm.py" "tests/test_gpu_tx.py::test_tx_am_modulator" -m gpu -q -x > "$OUT/1_racecheck.log" 2>&1
echo "racecheck exit $?" | tee -a "$OUT/summary.txt"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" "$OUT/1_racecheck.log" | tail -6 | tee -a "$OUT/summary.txt"
