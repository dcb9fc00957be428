#!/bin/bash
# round 2, GPU call v: slices per call (QRL_NSUB) for the two latency-bound configurations; QPSK without a private SM partition
set -u
OUT=gpurun_out/r02_v
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
(timeout 300 python tools/cfg_time.py
 for n in 1 2 4 6; do QRL_NSUB=$n timeout 300 python tools/cfg_time.py cfg2; done
 for n in 4 8 24; do QRL_NSUB=$n timeout 300 python tools/cfg_time.py qpsk; done
 QRL_QPSK_PARTITION=1 QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py qpsk) 2>&1 | grep -v Warning | tee "$OUT/1_ab.txt"
