#!/usr/bin/env python
"""Copies the artefacts of tools/gpu_r02_final2.sh from gpurun_out/r02_final2/ (scratch) into profiles/ under their committed names and
refreshes profiles/fir_traffic_bytes.json from the stage-1 FIR's `--set full` capture."""
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "r02_final2")
DST = os.path.join(ROOT, "profiles")
for a, b in (("3_bench.json", "r02_final2_bench.json"), ("2_bench_reference_arm.json", "r02_final2_bench_reference_arm.json"),
             ("4_launches_step.csv", "r02_final2_launches_step.csv"), ("5_ncu_full_summary.csv", "r02_final2_ncu_full_summary.csv"),
             ("summary.txt", "r02_final2_gpu_tier_summary.txt")):
    shutil.copy(os.path.join(SRC, a), os.path.join(DST, b))
rd = wr = dur = None
for rep, kern, metric, unit, value in list(csv.reader(open(os.path.join(SRC, "5_ncu_full_summary.csv"))))[1:]:
    if rep != "ncu_fir_poly.ncu-rep":
        continue
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit)
    if metric == "dram__bytes_read.sum":
        rd = float(value) * scale
    if metric == "dram__bytes_write.sum":
        wr = float(value) * scale
    if metric == "gpu__time_duration.sum":
        dur = float(value) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[unit]
n = 64 * (1 << 22)
json.dump({"per_sample": (rd + wr) / n, "dram_bytes_read": rd, "dram_bytes_write": wr, "samples_per_launch": n,
           "source": "ncu --set full --clock-control none, gpurun_out/r02_final2/ncu_fir_poly.ncu-rep (end of round 2), summary in profiles/r02_final2_ncu_full_summary.csv",
           "kernel": "fir_decim_poly_kernel<50,9,8,128,8>", "duration_us_under_ncu": dur},
          open(os.path.join(DST, "fir_traffic_bytes.json"), "w"), indent=1)
print("traffic per sample", (rd + wr) / n, "duration", dur)
