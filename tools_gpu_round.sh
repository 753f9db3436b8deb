#!/bin/bash
# Round 2, GPU session 3: 5x5 + sliced wide layers on the tensor cores (units, model parity, anime fps), PRMT A/B, wide-tile timeline.
O=gpurun_out/r2_s3
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=30 > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu: rc=$? $(tail -1 $O/pytest_gpu.log)"
RIFE_B200_WIDE=1 timeout 600 python -m pytest tests/test_tc_conv_gpu.py tests/test_parity_gpu.py -q -p no:cacheprovider -k "wide or v23 or every_model or fused or v46_plain or golden" > $O/pytest_wide1.log 2>&1
stamp "pytest with RIFE_B200_WIDE=1: rc=$? $(tail -1 $O/pytest_wide1.log)"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
timeout 300 $B > $O/bench_prmt.json 2> $O/bench_prmt.err
stamp "bench PRMT (default build) rc=$? $(cut -c1-120 $O/bench_prmt.json)"
RIFE_B200_LIB=$PWD/rife-ncnn-vulkan_b200/lib_lean/librife_b200.so timeout 300 $B > $O/bench_noprmt.json 2> $O/bench_noprmt.err
stamp "bench NO_PRMT build rc=$? $(cut -c1-120 $O/bench_noprmt.json)"
timeout 300 $B > $O/bench_prmt2.json 2> $O/bench_prmt2.err
stamp "bench PRMT again rc=$? $(cut -c1-120 $O/bench_prmt2.json)"
RIFE_B200_WIDE=0 timeout 120 python tools/timeline.py > $O/timeline_wide0.txt 2>&1
RIFE_B200_WIDE=1 timeout 120 python tools/timeline.py > $O/timeline_wide1.txt 2>&1
stamp "timelines done"
timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 2 > $O/anime_tta_fps.txt 2>&1
stamp "anime -x -z fps: $(tail -1 $O/anime_tta_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-anime --frames 8 > $O/anime_plain_fps.txt 2>&1
stamp "anime plain fps: $(tail -1 $O/anime_plain_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-v2.3 --frames 8 > $O/v23_plain_fps.txt 2>&1
stamp "v2.3 plain fps: $(tail -1 $O/v23_plain_fps.txt)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file $O/anime_plain_launches.csv \
    python tools/profile_model.py --model rife-anime --frames 1 --warmup 1 > $O/ncu_anime.log 2>&1
stamp "ncu anime launch list rc=$?"
python tools/summarise_launches.py $O/anime_plain_launches.csv 30 > $O/anime_plain_launches_summary.txt 2>&1
RIFE_BENCH_PAIRS=8 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file $O/launches_1080p.csv \
    python bench.py --only --no-cpu-baseline --no-process-leg --steps 1 --warmup 3 --lanes 1 > $O/ncu_launches.log 2>&1
stamp "ncu v4.6 launch list rc=$?"
python tools/summarise_launches.py $O/launches_1080p.csv 30 > $O/launches_1080p_summary.txt 2>&1
cat $O/summary.txt
