#!/bin/bash
# Round 2, GPU session 1: the suite, the new bench line (1080p + 4K + parity + e2e_process), HBM-kernel bandwidths + ncu,
# LEAN build A/B, precision of plain block-head tensors.
O=gpurun_out/r2_s1
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --maxfail=25 > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu: rc=$? $(tail -1 $O/pytest_gpu.log)"
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
stamp "bench.py rc=$? $(cut -c1-300 $O/bench.json)"
timeout 200 python tools/bench_hbm.py --size 1080p > $O/hbm_1080p.txt 2>&1
stamp "bench_hbm 1080p rc=$?"
timeout 200 python tools/bench_hbm.py --size 4k > $O/hbm_4k.txt 2>&1
stamp "bench_hbm 4k rc=$?"
RIFE_B200_LIB=$PWD/rife-ncnn-vulkan_b200/lib_lean/librife_b200.so timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg > $O/bench_lean.json 2> $O/bench_lean.err
stamp "bench LEAN rc=$? $(cut -c1-200 $O/bench_lean.json)"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg > $O/bench_nolean.json 2> $O/bench_nolean.err
stamp "bench default (same session, for the A/B) rc=$? $(cut -c1-200 $O/bench_nolean.json)"
timeout 300 python tools/precision_heads.py > $O/precision_heads.txt 2>&1
stamp "precision_heads rc=$?"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg --plain-blocks 252 > $O/bench_plainheads.json 2> $O/bench_plainheads.err
stamp "bench plain heads (mask 252) rc=$? $(cut -c1-200 $O/bench_plainheads.json)"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg --model rife-v4 --timestep 0.25 > $O/bench_v4.json 2> $O/bench_v4.err
stamp "bench rife-v4 rc=$? $(cut -c1-200 $O/bench_v4.json)"
timeout 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(preproc_kernel|postproc_tta_kernel|postproc_plain_kernel|flow_tta_avg_kernel|warp_kernel|temporal_merge)' -f -o $O/hbm_kernels \
    python tools/bench_hbm.py --size 1080p --ncu > $O/ncu_hbm.log 2>&1
stamp "ncu hbm kernels rc=$?"
timeout 60 ncu -i $O/hbm_kernels.ncu-rep --page details --csv > $O/hbm_kernels_details.csv 2>> $O/ncu_hbm.log
timeout 60 ncu -i $O/hbm_kernels.ncu-rep --page raw --csv > $O/hbm_kernels_raw.csv 2>> $O/ncu_hbm.log
stamp "exports done"
cat $O/summary.txt
