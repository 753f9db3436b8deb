#!/bin/bash
# Round 2, GPU session 4: the shipped state -- full suite, the default bench line, HBM kernels (float4 flow_tta_avg), config 4,
# strong-scaling mode at N = 1, ncu --set full of the dominant kernel in the step and of the 5x5 variant.
O=gpurun_out/r2_s4
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=30 > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu: rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
stamp "bench.py rc=$? $(cut -c1-200 $O/bench.json)"
timeout 200 python tools/bench_hbm.py --size 1080p > $O/hbm_1080p.txt 2>&1
timeout 200 python tools/bench_hbm.py --size 4k > $O/hbm_4k.txt 2>&1
stamp "bench_hbm done"
timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 3 > $O/anime_tta_fps.txt 2>&1
stamp "anime -x -z fps: $(tail -1 $O/anime_tta_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-anime --frames 8 > $O/anime_plain_fps.txt 2>&1
stamp "anime plain fps: $(tail -1 $O/anime_plain_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-v2.3 --frames 8 > $O/v23_plain_fps.txt 2>&1
stamp "v2.3 plain fps: $(tail -1 $O/v23_plain_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-v4.6 --tta --frames 4 > $O/v46_tta_fps.txt 2>&1
stamp "v4.6 -x fps: $(tail -1 $O/v46_tta_fps.txt)"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg --scaling strong --steps 3 > $O/bench_strong_n1.json 2> $O/bench_strong_n1.err
stamp "bench --scaling strong (N=1) rc=$? $(cut -c1-160 $O/bench_strong_n1.json)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file $O/anime_plain_launches.csv \
    python tools/profile_model.py --model rife-anime --frames 1 --warmup 1 > $O/ncu_anime.log 2>&1
python tools/summarise_launches.py $O/anime_plain_launches.csv 30 > $O/anime_plain_launches_summary.txt 2>&1
stamp "ncu anime launch list done"
RIFE_BENCH_PAIRS=8 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:tc_conv3x3_kernel<.int.64, .int.4, .int.3, .int.9, .int.0>' -s 60 -c 2 -f -o $O/conv64_step \
    python bench.py --only --no-cpu-baseline --no-process-leg --steps 1 --warmup 3 --lanes 1 > $O/ncu_conv64.log 2>&1
stamp "ncu conv64 in the step rc=$?"
timeout 60 ncu -i $O/conv64_step.ncu-rep --page details --csv > $O/conv64_step_details.csv 2>> $O/ncu_conv64.log
timeout 60 ncu -i $O/conv64_step.ncu-rep --page raw --csv > $O/conv64_step_raw.csv 2>> $O/ncu_conv64.log
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:tc_conv3x3_kernel<.int.(48|96), .int.[24], .int.2, .int.5' -s 4 -c 4 -f -o $O/conv5x5 \
    python tools/profile_model.py --model rife-anime --frames 1 --warmup 1 > $O/ncu_conv5x5.log 2>&1
stamp "ncu conv5x5 rc=$?"
timeout 60 ncu -i $O/conv5x5.ncu-rep --page details --csv > $O/conv5x5_details.csv 2>> $O/ncu_conv5x5.log
stamp "exports done"
cat $O/summary.txt
