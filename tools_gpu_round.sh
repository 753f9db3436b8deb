#!/bin/bash
# One GPU-box session: kernel self-tests, parity, bench, launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
(timeout 300 python -m pytest tests/test_tc_conv_gpu.py -m gpu -q --durations=5 -x 2>&1 | tail -40) > gpurun_out/tc_conv_tests.txt
tail -15 gpurun_out/tc_conv_tests.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --durations=30 2>&1 | tail -70) > gpurun_out/parity_tests.txt
tail -40 gpurun_out/parity_tests.txt
(timeout 300 python bench.py --steps 3 --warmup 3 2>&1 | tail -5) > gpurun_out/bench_1080p.txt
cat gpurun_out/bench_1080p.txt
