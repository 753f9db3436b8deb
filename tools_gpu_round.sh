#!/bin/bash
# Round 2, GPU session 11: spatial TTA dealt to the lanes (tta_fork / tta_join) -- parity of every TTA mode, fps vs one lane.
O=gpurun_out/r2_s11
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_hbm_kernels_gpu.py -q -p no:cacheprovider -k "tta or golden or bgr or uhd or hbm or flow or postproc or preproc" > $O/pytest_tta.log 2>&1
stamp "pytest TTA modes: rc=$? $(tail -1 $O/pytest_tta.log)"
for L in 1 2 4; do
  timeout 120 python tools/profile_model.py --model rife-anime --tta --tta-temporal --frames 3 --lanes $L > $O/anime_tta_l$L.txt 2>&1
  stamp "$(tail -1 $O/anime_tta_l$L.txt)"
done
for L in 1 4; do
  timeout 120 python tools/profile_model.py --model rife-v4.6 --tta --frames 6 --lanes $L > $O/v46_tta_l$L.txt 2>&1
  stamp "$(tail -1 $O/v46_tta_l$L.txt)"
  timeout 120 python tools/profile_model.py --model rife-v2.3 --tta --tta-temporal --frames 3 --lanes $L > $O/v23_tta_l$L.txt 2>&1
  stamp "$(tail -1 $O/v23_tta_l$L.txt)"
done
timeout 120 python tools/profile_model.py --model rife-anime --frames 8 > $O/anime_plain.txt 2>&1
stamp "$(tail -1 $O/anime_plain.txt)"
cat $O/summary.txt
