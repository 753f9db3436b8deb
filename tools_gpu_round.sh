mkdir -p gpurun_out/s6
RIFE_BENCH_PAIRS=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s6/launches.csv python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/s6/b.log 2>&1
tail -2 gpurun_out/s6/b.log | cut -c1-300
wc -l gpurun_out/s6/launches.csv
