#!/bin/bash
# One GPU-box session: ncu launch list + full capture of the tensor-core conv, 4K bench, CPU thread sweep.
mkdir -p gpurun_out
export RIFE_BENCH_PAIRS=1
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_1080p.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_stdout.txt 2>&1)
wc -l gpurun_out/launches_1080p.csv
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv3x3 -s 10 -c 2 -o gpurun_out/prof_tc_conv64 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_stdout.txt 2>&1)
ls -la gpurun_out/
unset RIFE_BENCH_PAIRS
(timeout 300 python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench_4k.txt
cat gpurun_out/bench_4k.txt
for t in 16 32 64 128; do
  python - <<PY
import sys; sys.path.insert(0,'tests'); import parity
a,b=parity.synth.pair(1920,1080)
_,i=parity.run_oracle("rife-v4.6",a,b,threads=$t,repeat=2,warmup=1)
print("threads",$t,i)
PY
done > gpurun_out/cpu_thread_sweep.txt 2>&1
cat gpurun_out/cpu_thread_sweep.txt
