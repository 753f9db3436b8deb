#!/bin/bash
mkdir -p gpurun_out
for m in 0 8 12 15; do
(timeout 300 python bench.py --steps 5 --warmup 3 --plain-blocks $m --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_1080p_pm$m.txt
python -c "import json;d=json.load(open('gpurun_out/bench_1080p_pm$m.txt'));print('1080p plain mask',$m,round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done
for m in 0 12 15; do
(timeout 300 python bench.py --steps 3 --warmup 3 --workload 4k --plain-blocks $m --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_4k_pm$m.txt
python -c "import json;d=json.load(open('gpurun_out/bench_4k_pm$m.txt'));print('4k plain mask',$m,round(d['value'],1),round(d['e2e']['value'],1),d['clocks']['sm_mhz'])"
done
