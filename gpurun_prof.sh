mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_b.json 2> gpurun_out/bench_b_err.log
tail -2 gpurun_out/bench_b_err.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_b_ref.json 2>> gpurun_out/bench_b_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_r01_b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b1.log 2>&1
QRL_NSUB=1 ncu --set full --clock-control none --import-source on -k regex:"fir_decim_poly|symsync_kernel|viterbi_k7" -s 9 -c 3 -o gpurun_out/top3_r01_b python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_b2.log 2>&1
tail -2 gpurun_out/ncu_b2.log
ls -la gpurun_out
