#!/bin/bash
mkdir -p gpurun_out
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_1080p_plain_b8 python tools/profile_tc.py 1080p 0 8 > gpurun_out/ncu_tc_stdout.txt 2>&1)
(timeout 300 ncu --set full --clock-control none -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_4k_plain_b2 python tools/profile_tc.py 4k 0 2 >> gpurun_out/ncu_tc_stdout.txt 2>&1)
# launch list of one batched step (8 pairs, 1 lane) at 1080p
RIFE_BENCH_PAIRS=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_1080p_batched.csv python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ncu_bench_stdout.txt 2>&1
wc -l gpurun_out/launches_1080p_batched.csv
ls -la gpurun_out/*.ncu-rep
