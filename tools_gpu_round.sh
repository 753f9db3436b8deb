#!/bin/bash
# Round 2, GPU session 7: knock-out study of the conv mainloop (events, no in-kernel instrumentation), per-image scheduling probe.
O=gpurun_out/r2_s7
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 300 python tools/knockout.py > $O/knockout.txt 2>&1
stamp "knockout rc=$?"
RIFE_B200_WIDE=1 timeout 300 python tools/knockout.py > $O/knockout_wide1.txt 2>&1
stamp "knockout wide rc=$?"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
timeout 300 $B --lanes 1 > $O/bench_b8_l1.json 2> $O/bench_b8_l1.err
stamp "bench batch auto lanes 1 rc=$? $(cut -c1-110 $O/bench_b8_l1.json)"
timeout 300 $B --lanes 1 --batch 1 > $O/bench_b1_l1.json 2> $O/bench_b1_l1.err
stamp "bench batch 1 lanes 1 rc=$? $(cut -c1-110 $O/bench_b1_l1.json)"
timeout 300 $B --lanes 1 --batch 2 > $O/bench_b2_l1.json 2> $O/bench_b2_l1.err
stamp "bench batch 2 lanes 1 rc=$? $(cut -c1-110 $O/bench_b2_l1.json)"
timeout 300 $B --lanes 3 > $O/bench_l3.json 2> $O/bench_l3.err
stamp "bench lanes 3 rc=$? $(cut -c1-110 $O/bench_l3.json)"
timeout 300 python bench.py --only --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
stamp "bench default (KS=2 default, host link) rc=$? $(cut -c1-110 $O/bench_default.json)"
cat $O/summary.txt
