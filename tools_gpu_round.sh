#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_tc_conv_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/tc_conv_tests.txt
tail -3 gpurun_out/tc_conv_tests.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/parity_tests.txt
tail -6 gpurun_out/parity_tests.txt
for cfg in "2 1" "2 2" "2 4" "2 8" "3 4" "1 8" "3 0"; do
set -- $cfg
(timeout 300 python bench.py --steps 5 --warmup 3 --lanes $1 --batch $2 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_1080p_l$1_b$2.txt
python -c "import json;d=json.load(open('gpurun_out/bench_1080p_l$1_b$2.txt'));print('1080p lanes',$1,'batch',$2,round(d['value'],1),round(d['e2e']['value'],1),d['clocks'])"
done
for cfg in "2 1" "2 2" "3 1"; do
set -- $cfg
(timeout 300 python bench.py --steps 3 --warmup 3 --workload 4k --lanes $1 --batch $2 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_4k_l$1_b$2.txt
python -c "import json;d=json.load(open('gpurun_out/bench_4k_l$1_b$2.txt'));print('4k lanes',$1,'batch',$2,round(d['value'],1),round(d['e2e']['value'],1))"
done
