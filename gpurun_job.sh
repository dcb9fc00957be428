mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fir_|symsync|viterbi|qdemod|hist_update" -s 40 -c 80 --csv --log-file gpurun_out/launches_r01_c.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log | cut -c1-200
