#!/bin/bash
# Round 2, GPU session 6: two 16-channel chunks per pipeline stage (RIFE_B200_KS=2) A/B, 8-channel heads on the tensor path, D2H on two streams.
O=gpurun_out/r2_s6
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
RIFE_B200_KS=2 timeout 600 python -m pytest tests/test_tc_conv_gpu.py tests/test_parity_gpu.py -q -p no:cacheprovider -k "conv3x3 or v46 or v4_ or every_model or golden or fused" > $O/pytest_ks2.log 2>&1
stamp "pytest with RIFE_B200_KS=2: rc=$? $(tail -1 $O/pytest_ks2.log)"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_tc_conv_gpu.py -q -p no:cacheprovider -k "every_model or tta or golden or uhd or anime or conv" > $O/pytest_models.log 2>&1
stamp "pytest models (8-channel heads on tc): rc=$? $(tail -1 $O/pytest_models.log)"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
for i in 1 2; do
  RIFE_B200_KS=1 timeout 300 $B > $O/bench_ks1_$i.json 2> $O/bench_ks1_$i.err
  stamp "bench KS=1 #$i rc=$? $(cut -c1-110 $O/bench_ks1_$i.json)"
  RIFE_B200_KS=2 timeout 300 $B > $O/bench_ks2_$i.json 2> $O/bench_ks2_$i.err
  stamp "bench KS=2 #$i rc=$? $(cut -c1-110 $O/bench_ks2_$i.json)"
done
RIFE_B200_KS=2 RIFE_B200_WIDE=1 timeout 300 $B > $O/bench_ks2_wide1.json 2> $O/bench_ks2_wide1.err
stamp "bench KS=2 WIDE=1 rc=$? $(cut -c1-110 $O/bench_ks2_wide1.json)"
RIFE_B200_KS=2 timeout 300 $B --workload 4k > $O/bench_4k_ks2.json 2> $O/bench_4k_ks2.err
stamp "bench 4K KS=2 rc=$? $(cut -c1-110 $O/bench_4k_ks2.json)"
RIFE_B200_KS=1 timeout 120 python tools/timeline.py > $O/timeline_ks1.txt 2>&1
RIFE_B200_KS=2 timeout 120 python tools/timeline.py > $O/timeline_ks2.txt 2>&1
RIFE_B200_KS=2 RIFE_B200_WIDE=1 timeout 120 python tools/timeline.py > $O/timeline_ks2_wide1.txt 2>&1
stamp "timelines done"
timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 3 > $O/anime_tta_fps.txt 2>&1
stamp "anime -x -z fps: $(tail -1 $O/anime_tta_fps.txt)"
timeout 120 python tools/profile_model.py --model rife-anime --frames 8 > $O/anime_plain_fps.txt 2>&1
stamp "anime plain fps: $(tail -1 $O/anime_plain_fps.txt)"
cat $O/summary.txt
