#!/bin/bash
# round 2, GPU call y: 4FSK chain with 256 channels -- which stage bounds it, and does the Viterbi decoder belong on the wide partition there
set -u
OUT=gpurun_out/r02_y
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
(timeout 300 python tools/cfg_time.py cfg2:256
 QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py cfg2:256
 QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py cfg2:64
 QRL_LOOP_SMS=8 QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py cfg2:256
 QRL_NSUB=4 QRL_FEC_ON_PAR=1 timeout 300 python tools/cfg_time.py cfg2:256) 2>&1 | grep -v Warning | tee "$OUT/1_ab.txt"
