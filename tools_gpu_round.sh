#!/bin/bash
# Round 2, GPU session 2: wide-tile conv + packed heads (unit tests, parity, A/B), vectorised HBM kernels, anime launch list.
O=gpurun_out/r2_s2
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 600 python -m pytest tests/test_tc_conv_gpu.py tests/test_hbm_kernels_gpu.py -q -p no:cacheprovider > $O/pytest_units.log 2>&1
stamp "pytest units (tc conv incl. wide, hbm kernels): rc=$? $(tail -1 $O/pytest_units.log)"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -p no:cacheprovider --maxfail=30 > $O/pytest_parity.log 2>&1
stamp "pytest parity: rc=$? $(tail -1 $O/pytest_parity.log)"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
timeout 300 $B > $O/bench_wide1_pack0.json 2> $O/bench_wide1_pack0.err
stamp "bench wide=1 pack=0 rc=$? $(cut -c1-120 $O/bench_wide1_pack0.json)"
RIFE_B200_WIDE=0 timeout 300 $B > $O/bench_wide0_pack0.json 2> $O/bench_wide0_pack0.err
stamp "bench wide=0 pack=0 rc=$? $(cut -c1-120 $O/bench_wide0_pack0.json)"
timeout 300 $B --head-pack 1 > $O/bench_wide1_pack1.json 2> $O/bench_wide1_pack1.err
stamp "bench wide=1 pack=1 rc=$? $(cut -c1-120 $O/bench_wide1_pack1.json)"
timeout 300 $B --head-pack 1 --workload 4k > $O/bench_4k_wide1_pack1.json 2> $O/bench_4k_wide1_pack1.err
stamp "bench 4k wide=1 pack=1 rc=$? $(cut -c1-120 $O/bench_4k_wide1_pack1.json)"
timeout 200 python tools/bench_hbm.py --size 1080p > $O/hbm_1080p.txt 2>&1
stamp "bench_hbm 1080p rc=$?"
timeout 200 python tools/bench_hbm.py --size 4k > $O/hbm_4k.txt 2>&1
stamp "bench_hbm 4k rc=$?"
RIFE_BENCH_PAIRS=8 timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:tc_conv3x3_kernel<.int.64, .int.4, .int.3, .int.9, .int.1>' -s 16 -c 2 -f -o $O/conv64_wide \
    python bench.py --only --no-cpu-baseline --no-process-leg --steps 1 --warmup 3 --lanes 1 > $O/ncu_conv64.log 2>&1
stamp "ncu wide conv rc=$?"
timeout 60 ncu -i $O/conv64_wide.ncu-rep --page details --csv > $O/conv64_wide_details.csv 2>> $O/ncu_conv64.log
timeout 60 ncu -i $O/conv64_wide.ncu-rep --page raw --csv > $O/conv64_wide_raw.csv 2>> $O/ncu_conv64.log
timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 2 > $O/anime_tta_fps.txt 2>&1
stamp "anime -x -z fps: $(tail -1 $O/anime_tta_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-anime --frames 4 > $O/anime_plain_fps.txt 2>&1
stamp "anime plain fps: $(tail -1 $O/anime_plain_fps.txt)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file $O/anime_plain_launches.csv \
    python tools/profile_model.py --model rife-anime --frames 1 --warmup 1 > $O/ncu_anime.log 2>&1
stamp "ncu anime launch list rc=$?"
python tools/summarise_launches.py $O/anime_plain_launches.csv 30 > $O/anime_plain_launches_summary.txt 2>&1
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(preproc_kernel|postproc_tta_kernel|postproc_plain_kernel|flow_tta_avg_kernel|warp_kernel|temporal_merge)' -f -o $O/hbm_kernels \
    python tools/bench_hbm.py --size 1080p --ncu > $O/ncu_hbm.log 2>&1
stamp "ncu hbm kernels rc=$?"
timeout 60 ncu -i $O/hbm_kernels.ncu-rep --page details --csv > $O/hbm_kernels_details.csv 2>> $O/ncu_hbm.log
stamp "exports done"
cat $O/summary.txt
