#!/bin/bash
# Round 2, GPU session 13: the shipped state -- smoke(), full suite, the default bench line, config 4 / 5 side measurements.
O=gpurun_out/r2_s13
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
stamp "smoke(): rc=$? $(tail -1 $O/smoke.log)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=30 > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu: rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
stamp "bench.py rc=$? $(cut -c1-200 $O/bench.json)"
timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 3 > $O/anime_tta_fps.txt 2>&1
stamp "$(tail -1 $O/anime_tta_fps.txt)"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg --model rife-v4 --timestep 0.25 > $O/bench_v4.json 2> $O/bench_v4.err
stamp "bench rife-v4 rc=$? $(cut -c1-120 $O/bench_v4.json)"
RIFE_BENCH_PAIRS=8 timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv --log-file $O/launches_1080p.csv \
    python bench.py --only --no-cpu-baseline --no-process-leg --steps 2 --warmup 3 --lanes 1 > $O/ncu_launches.log 2>&1
python tools/summarise_launches.py $O/launches_1080p.csv 30 > $O/launches_1080p_summary.txt 2>&1
stamp "ncu launch list of the bench command done"
cat $O/summary.txt
