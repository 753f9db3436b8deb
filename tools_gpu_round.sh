#!/bin/bash
# One gpurun session: validates the two opt-in changes of this round (recompute_fm, paired MMA issue) and measures them
# against the current default.  Everything lands in gpurun_out/s20/; every step has its own timeout.
O=gpurun_out/s20
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > $O/gpu.txt 2>&1

# A. tensor-core unit tests under each kernel variant (separate processes: a faulting variant cannot poison the others)
for PM in 1 2 3; do
  RIFE_B200_PAIR=$PM timeout 150 python -m pytest tests/test_tc_conv_gpu.py -q -m gpu -p no:cacheprovider > $O/tc_unit_pair$PM.log 2>&1
  stamp "tc unit tests PAIR=$PM rc=$? $(tail -1 $O/tc_unit_pair$PM.log)"
done

# B. the whole GPU suite with everything switched on
RIFE_B200_PAIR=3 RIFE_B200_RECOMPUTE_FM=2 timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_all_on.log 2>&1
stamp "pytest -m gpu, PAIR=3 RECOMPUTE_FM=2: rc=$? $(tail -1 $O/pytest_all_on.log)"

# C. 1080p A/B (10 timed steps of 128 pairs each, as the driver runs it)
bench() {  # name, env..., -- args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 240 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err
  local rc=$?
  stamp "bench $name rc=$rc $(python - <<EOF
import json
try:
    d = json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1])
    print("value %.0f e2e %.0f ms %.2f clk %s conv_us %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print("no result:", e)
EOF
)"
}
bench 1080p_default X=1 -- --steps 10 --warmup 3
bench 1080p_rc2 X=1 -- --steps 10 --warmup 3 --recompute-fm 2 --no-cpu-baseline
bench 1080p_pair3_rc2 RIFE_B200_PAIR=3 -- --steps 10 --warmup 3 --recompute-fm 2 --no-cpu-baseline
bench 1080p_pair1 RIFE_B200_PAIR=1 -- --steps 10 --warmup 3 --no-cpu-baseline
bench 1080p_rc1 X=1 -- --steps 10 --warmup 3 --recompute-fm 1 --no-cpu-baseline
bench 1080p_pair1_rc2 RIFE_B200_PAIR=1 -- --steps 10 --warmup 3 --recompute-fm 2 --no-cpu-baseline

# D. the whole GPU suite with the defaults of this commit, then smoke
timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_default.log 2>&1
stamp "pytest -m gpu, defaults: rc=$? $(tail -1 $O/pytest_default.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
stamp "smoke rc=$? $(tail -1 $O/smoke.log)"

# E. 4K A/B
bench 4k_default X=1 -- --steps 10 --warmup 3 --workload 4k --no-cpu-baseline
bench 4k_pair3_rc2 RIFE_B200_PAIR=3 -- --steps 10 --warmup 3 --workload 4k --recompute-fm 2 --no-cpu-baseline

# F. ncu: full capture of the dominant kernel as launched in the timed step (8 x 480x272, plain fp16), both issue forms
for PM in 0 3; do
  RIFE_B200_PAIR=$PM timeout 200 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3 -s 2 -c 1 -f -o $O/conv64_1080p_b8_plain_pair$PM \
      python tools/profile_tc.py 1080p 0 8 > $O/ncu_conv_pair$PM.log 2>&1
  timeout 60 ncu -i $O/conv64_1080p_b8_plain_pair$PM.ncu-rep --page details --csv > $O/conv64_1080p_b8_plain_pair${PM}_details.csv 2>> $O/ncu_conv_pair$PM.log
  stamp "ncu conv PAIR=$PM rc=$?"
done
# launch list of one 8-pair batch, one lane
RIFE_BENCH_PAIRS=8 RIFE_B200_PAIR=3 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_1080p_pair3_rc2.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --recompute-fm 2 --no-cpu-baseline > $O/ncu_launches_on.log 2>&1
stamp "ncu launch list (all on) rc=$?"
RIFE_BENCH_PAIRS=8 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_1080p_default.csv \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > $O/ncu_launches_default.log 2>&1
stamp "ncu launch list (default) rc=$?"

# G. per-stage CUDA-event times of one lane (diagnostics)
RIFE_B200_KTIME=1 RIFE_BENCH_PAIRS=16 timeout 120 python bench.py --steps 3 --warmup 3 --lanes 1 --no-cpu-baseline > $O/ktime_default.json 2> $O/ktime_default.txt
RIFE_B200_KTIME=1 RIFE_BENCH_PAIRS=16 RIFE_B200_PAIR=3 timeout 120 python bench.py --steps 3 --warmup 3 --lanes 1 --recompute-fm 2 --no-cpu-baseline > $O/ktime_all_on.json 2> $O/ktime_all_on.txt
stamp "ktime done"
cat $O/summary.txt
# H. mainloop cost per 16-channel stage (empty epilogue) for each issue form
for PM in 0 1 3; do
  RIFE_B200_PAIR=$PM timeout 100 python tools/mma_cost.py > $O/mma_cost_pair$PM.txt 2>&1
done
stamp "mma_cost done"
