#!/bin/bash
# Round 2, GPU session 8: reverse tile order on alternate convolutions (RIFE_B200_SNAKE), K-loop rotation for streamed weights (RIFE_B200_KROT).
O=gpurun_out/r2_s8
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
RIFE_B200_SNAKE=1 RIFE_B200_KROT=1 timeout 600 python -m pytest tests/test_tc_conv_gpu.py tests/test_parity_gpu.py -q -p no:cacheprovider -k "conv or v46 or v4_ or every_model or golden or fused or batched" > $O/pytest_snake_krot.log 2>&1
stamp "pytest with SNAKE=1 KROT=1: rc=$? $(tail -1 $O/pytest_snake_krot.log)"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
for i in 1 2; do
  timeout 300 $B > $O/bench_base_$i.json 2> $O/bench_base_$i.err
  stamp "bench base #$i rc=$? $(cut -c1-110 $O/bench_base_$i.json)"
  RIFE_B200_SNAKE=1 timeout 300 $B > $O/bench_snake_$i.json 2> $O/bench_snake_$i.err
  stamp "bench SNAKE #$i rc=$? $(cut -c1-110 $O/bench_snake_$i.json)"
  RIFE_B200_KROT=1 timeout 300 $B > $O/bench_krot_$i.json 2> $O/bench_krot_$i.err
  stamp "bench KROT #$i rc=$? $(cut -c1-110 $O/bench_krot_$i.json)"
done
RIFE_B200_SNAKE=1 RIFE_B200_KROT=1 timeout 300 $B > $O/bench_both.json 2> $O/bench_both.err
stamp "bench SNAKE+KROT rc=$? $(cut -c1-110 $O/bench_both.json)"
RIFE_B200_SNAKE=1 RIFE_B200_KROT=1 timeout 300 $B --workload 4k > $O/bench_both_4k.json 2> $O/bench_both_4k.err
stamp "bench SNAKE+KROT 4K rc=$? $(cut -c1-110 $O/bench_both_4k.json)"
timeout 300 $B --workload 4k > $O/bench_base_4k.json 2> $O/bench_base_4k.err
stamp "bench base 4K rc=$? $(cut -c1-110 $O/bench_base_4k.json)"
RIFE_B200_KROT=1 timeout 300 python tools/knockout.py > $O/knockout_krot.txt 2>&1
stamp "knockout with KROT rc=$?"
cat $O/summary.txt
