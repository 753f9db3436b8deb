#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_tc_conv_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/tc_conv_tests.txt
tail -5 gpurun_out/tc_conv_tests.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/parity_tests.txt
tail -8 gpurun_out/parity_tests.txt
for lanes in 1 2 3; do
(timeout 300 python bench.py --steps 5 --warmup 3 --lanes $lanes --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench_1080p_l$lanes.txt
cat gpurun_out/bench_1080p_l$lanes.txt
done
(timeout 300 python bench.py --steps 5 --warmup 3 --workload 4k --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench_4k.txt
cat gpurun_out/bench_4k.txt
export RIFE_BENCH_PAIRS=1
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_1080p.csv \
   python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench_stdout.txt 2>&1)
wc -l gpurun_out/launches_1080p.csv
