#!/bin/bash
# Final validation session of round 1 (full snapshot: every model family, the reference CLI built against the shim).
O=gpurun_out/s21
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/gpu.txt 2>&1

timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu (defaults: paired issue on): rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
stamp "smoke rc=$? $(tail -1 $O/smoke.log)"

bench() {
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 240 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err
  local rc=$?
  stamp "bench $name rc=$rc $(python - <<EOF2
import json
try:
    d = json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1])
    print("value %.0f e2e %.0f ms %.2f clk %s conv_us %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print("no result:", e)
EOF2
)"
}
bench 1080p X=1 -- --steps 10 --warmup 3
bench 4k X=1 -- --steps 10 --warmup 3 --workload 4k --no-cpu-baseline

# option "combine" as the default: v4.6 / API tests again, then the threaded process() throughput with and without it
RIFE_B200_COMBINE=1 timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider -k "v46 or concurrent or batch or combined or golden or cli or edges" > $O/pytest_combine.log 2>&1
stamp "pytest, RIFE_B200_COMBINE=1 subset: rc=$? $(tail -1 $O/pytest_combine.log)"
timeout 200 python tools/bench_process_threads.py 1080p 24 > $O/process_threads_1080p.txt 2>&1
stamp "process() threads: $(grep -c calls/s $O/process_threads_1080p.txt) lines"

# ncu: launch list of one 8-pair batch (one lane), then full captures of the HBM-side outliers of block 3
RIFE_BENCH_PAIRS=8 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_1080p.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
stamp "ncu launch list rc=$?"
RIFE_BENCH_PAIRS=8 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(tc_conv3x3_kernel<32, 4, 4, 4>|head_update_kernel<1, 2, 2, 8|head_update_kernel<2, 4, 1, 8|tail_kernel)' -s 10 -c 10 -f -o $O/block3_hbm \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > $O/ncu_block3.log 2>&1
timeout 60 ncu -i $O/block3_hbm.ncu-rep --page details --csv > $O/block3_hbm_details.csv 2>> $O/ncu_block3.log
stamp "ncu block-3 captures rc=$?"
cat $O/summary.txt
