#!/bin/bash
O=gpurun_out/r2_s22
mkdir -p $O
timeout 300 python -m pytest tests/test_hbm_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "tta or v23 or anime or uhd or golden" 2>&1 | tail -3
timeout 300 python tools/bench_hbm.py --size 1080p > $O/hbm_1080p.txt 2> $O/hbm_1080p.err; grep -v "^{" $O/hbm_1080p.txt | cut -c1-260
timeout 300 python tools/bench_hbm.py --size 4k > $O/hbm_4k.txt 2> $O/hbm_4k.err; grep -v "^{" $O/hbm_4k.txt | cut -c1-260
