#!/bin/bash
# Round 2, GPU session 14: where the results' device -> host copies are queued (RIFE_B200_D2H) -- e2e A/B at 1080p and 4K.
O=gpurun_out/r2_s14
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
B="python bench.py --only --no-cpu-baseline --no-process-leg"
show() { python -c "import json,sys; d=json.load(open(sys.argv[1])); print('value %.0f e2e %.0f link %s numa %s' % (d['value'], d['e2e']['value'], d['config'].get('host_link_GBps'), d['config'].get('host_numa')))" $1; }
for i in 1 2 3; do
  RIFE_B200_D2H=0 timeout 300 $B > $O/bench_d2h0_$i.json 2> $O/bench_d2h0_$i.err; stamp "D2H=0 #$i $(show $O/bench_d2h0_$i.json)"
  RIFE_B200_D2H=1 timeout 300 $B > $O/bench_d2h1_$i.json 2> $O/bench_d2h1_$i.err; stamp "D2H=1 #$i $(show $O/bench_d2h1_$i.json)"
done
RIFE_B200_D2H=0 timeout 300 $B --workload 4k > $O/bench4k_d2h0.json 2> $O/bench4k_d2h0.err; stamp "4K D2H=0 $(show $O/bench4k_d2h0.json)"
RIFE_B200_D2H=1 timeout 300 $B --workload 4k > $O/bench4k_d2h1.json 2> $O/bench4k_d2h1.err; stamp "4K D2H=1 $(show $O/bench4k_d2h1.json)"
RIFE_B200_D2H=1 timeout 300 $B --lanes 3 > $O/bench_d2h1_l3.json 2> $O/bench_d2h1_l3.err; stamp "D2H=1 lanes 3 $(show $O/bench_d2h1_l3.json)"
cat $O/summary.txt
