#!/bin/bash
# Round 2, GPU session 10: whole-chunk MMA issue blocks (RIFE_B200_CHUNK_ISSUE) -- units, parity, timeline, A/B.
O=gpurun_out/r2_s10
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 300 python -m pytest tests/test_tc_conv_gpu.py -q -p no:cacheprovider > $O/pytest_units.log 2>&1
stamp "pytest tc units: rc=$? $(tail -1 $O/pytest_units.log)"
timeout 200 python tools/knock_timeline.py > $O/knock_timeline_chunk1.txt 2>&1
RIFE_B200_CHUNK_ISSUE=0 timeout 200 python tools/knock_timeline.py > $O/knock_timeline_chunk0.txt 2>&1
stamp "timelines done: $(head -1 $O/knock_timeline_chunk1.txt)"
B="python bench.py --only --no-cpu-baseline --no-process-leg"
for i in 1 2; do
  RIFE_B200_CHUNK_ISSUE=0 timeout 300 $B > $O/bench_chunk0_$i.json 2> $O/bench_chunk0_$i.err
  stamp "bench CHUNK_ISSUE=0 #$i rc=$? $(cut -c1-110 $O/bench_chunk0_$i.json)"
  timeout 300 $B > $O/bench_chunk1_$i.json 2> $O/bench_chunk1_$i.err
  stamp "bench CHUNK_ISSUE=1 #$i rc=$? $(cut -c1-110 $O/bench_chunk1_$i.json)"
done
timeout 600 python -m pytest tests/test_parity_gpu.py -q -p no:cacheprovider -k "v46 or v4_ or every_model or golden or fused or batched or recompute" > $O/pytest_parity.log 2>&1
stamp "pytest parity subset: rc=$? $(tail -1 $O/pytest_parity.log)"
cat $O/summary.txt
