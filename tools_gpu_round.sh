#!/bin/bash
mkdir -p gpurun_out
# (1) launch list of the fused path: 3 frames at 1080p; the last frame's launches are the steady state
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_1080p_fused.csv python tools/profile_frame.py 1080p > gpurun_out/ncu_frame_stdout.txt 2>&1)
wc -l gpurun_out/launches_1080p_fused.csv
# (2) full capture of the dominant tensor-core kernel, alone
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_1080p python tools/profile_tc.py 1080p > gpurun_out/ncu_tc_stdout.txt 2>&1)
(timeout 300 ncu --set full --clock-control none -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_4k python tools/profile_tc.py 4k >> gpurun_out/ncu_tc_stdout.txt 2>&1)
# (3) full capture of the HBM kernels of the last frame (head x4, update x3, tail, preproc x2 = 10 per frame; skip 2 frames + self-check)
(timeout 400 ncu --set full --clock-control none -k regex:"head|update_kernel|tail_kernel|preproc" -s 20 -c 10 -o gpurun_out/prof_hbm_1080p python tools/profile_frame.py 1080p > gpurun_out/ncu_hbm_stdout.txt 2>&1)
ls -la gpurun_out | head -30
