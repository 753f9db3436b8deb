#!/bin/bash
O=gpurun_out/r2_s12
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
for L in 4 6 8; do
  timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 3 --lanes $L > $O/anime_tta_l$L.txt 2>&1
  stamp "$(tail -1 $O/anime_tta_l$L.txt)"
done
timeout 120 python tools/profile_model.py --model rife-v4.6 --tta --frames 6 --lanes 8 > $O/v46_tta_l8.txt 2>&1
stamp "$(tail -1 $O/v46_tta_l8.txt)"
nvidia-smi --query-gpu=memory.used --format=csv > $O/mem.txt
cat $O/summary.txt
