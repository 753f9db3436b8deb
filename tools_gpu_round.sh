#!/bin/bash
# Last session of round 1: the suite with option "combine" on by default, ncu captures of the block-3 HBM-side kernels.
O=gpurun_out/s22
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 300 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/pytest_gpu.log 2>&1
stamp "pytest -m gpu (final defaults): rc=$? $(tail -1 $O/pytest_gpu.log)"
RIFE_BENCH_PAIRS=8 timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:(tc_conv3x3_kernel<.int.32, .int.4, .int.4, .int.4>|head_update_kernel<.int.1, .int.2, .int.2|head_update_kernel<.int.2, .int.4|tail_kernel)' -s 10 -c 10 -f -o $O/block3_hbm \
    python bench.py --steps 1 --warmup 3 --lanes 1 --no-cpu-baseline > $O/ncu_block3.log 2>&1
stamp "ncu block-3 captures rc=$?"
timeout 60 ncu -i $O/block3_hbm.ncu-rep --page details --csv > $O/block3_hbm_details.csv 2>> $O/ncu_block3.log
timeout 60 ncu -i $O/block3_hbm.ncu-rep --page raw --csv > $O/block3_hbm_raw.csv 2>> $O/ncu_block3.log
stamp "exports done"
cat $O/summary.txt
