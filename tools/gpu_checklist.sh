#!/bin/bash
# One GPU-box call that confirms everything round 1 left unconfirmed, then re-measures.  Run from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_checklist.sh'
# Output lands in gpurun_out/checklist/ (merged back by gpurun).  Stages are independent: a failing stage does not stop the next.
set -u
OUT=gpurun_out/checklist
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1

# 1. the two chains written without GPU time (DMR receive, M17 modulator): memcheck first -- they have never executed
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 \
    python -m pytest tests/test_gpu_dmr.py tests/test_gpu_m17_tx.py -m gpu -x -q > "$OUT/1_unverified_memcheck.log" 2>&1
echo "unverified (memcheck) exit $?" | tee "$OUT/summary.txt"

# 1b. the kernels that gained a barrier after the last GPU run: hardware racecheck on one small case each
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_golden.py::test_golden_rx_cuda[ssb_usb]" \
    "tests/test_golden.py::test_golden_rx_cuda[nbfm_2500]" "tests/test_golden.py::test_golden_rx_cuda[bpsk_2k]" "tests/test_golden.py::test_golden_rx_cuda[m17]" \
    -m gpu -x -q > "$OUT/1b_racecheck.log" 2>&1
echo "racecheck exit $?" | tee -a "$OUT/summary.txt"

# 2. the same without the sanitizer (parity verdict at full speed)
timeout 300 python -m pytest tests/test_gpu_dmr.py tests/test_gpu_m17_tx.py -m gpu -q > "$OUT/2_unverified.log" 2>&1
echo "unverified exit $?" | tee -a "$OUT/summary.txt"

# 3. the whole GPU tier as the driver runs it
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/3_gpu_tier.log" 2>&1
echo "gpu tier exit $?" | tee -a "$OUT/summary.txt"

# 4. smoke + bench (both arms), as the driver runs them
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/4_smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/5_bench_reference.json" 2> "$OUT/5_bench_reference.err"
timeout 600 python bench.py --steps 10 --warmup 3 > "$OUT/5_bench.json" 2> "$OUT/5_bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
tail -c 600 "$OUT/5_bench.json" | tee -a "$OUT/summary.txt"

# 5. launch list of one short bench run (cold, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/6_launches.csv" \
    python bench.py --steps 2 --warmup 1 > "$OUT/6_bench_under_ncu.log" 2>&1
echo "ncu exit $?" | tee -a "$OUT/summary.txt"
