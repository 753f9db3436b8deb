#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -4) > gpurun_out/parity_tests.txt
tail -2 gpurun_out/parity_tests.txt
(timeout 900 python tools/precision_study.py 2>&1 | tail -40) > gpurun_out/precision_study.txt
grep -E '"plain_blocks": (12|15)' gpurun_out/precision_study.txt
(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_1080p_default.txt
python -c "import json;d=json.load(open('gpurun_out/bench_1080p_default.txt'));print('1080p',round(d['value'],1),round(d['e2e']['value'],1))"
(timeout 300 python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_4k_default.txt
python -c "import json;d=json.load(open('gpurun_out/bench_4k_default.txt'));print('4k',round(d['value'],1),round(d['e2e']['value'],1))"
