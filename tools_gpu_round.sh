#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -25) > gpurun_out/all_gpu_tests.txt
tail -14 gpurun_out/all_gpu_tests.txt
for lanes in 3 4; do
(timeout 300 python bench.py --steps 5 --warmup 3 --lanes $lanes --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_1080p_l$lanes.txt
python -c "import json;d=json.load(open('gpurun_out/bench_1080p_l$lanes.txt'));print('lanes',$lanes,d['value'],d['e2e']['value'])"
done
(timeout 300 python bench.py --steps 5 --warmup 3 --lanes 3 --workload 4k --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_4k_l3.txt
python -c "import json;d=json.load(open('gpurun_out/bench_4k_l3.txt'));print('4k lanes 3',d['value'],d['e2e']['value'])"
(timeout 300 python bench.py --steps 3 --warmup 3 --model rife-v4 --timestep 0.25 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_v4_1080p.txt
python -c "import json;d=json.load(open('gpurun_out/bench_v4_1080p.txt'));print('rife-v4 1080p t=0.25',d['value'],d['e2e']['value'])"
RIFE_BENCH_PAIRS=2 timeout 600 python bench.py --steps 3 --warmup 3 --model rife-anime --tta --tta-temporal --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_anime_tta.txt
python -c "import json;d=json.load(open('gpurun_out/bench_anime_tta.txt'));print('anime -x -z 1080p',d['value'],d['e2e']['value'])"
(timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1) > gpurun_out/bench_1080p_default.txt
cat gpurun_out/bench_1080p_default.txt
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1) > gpurun_out/bench_reference.txt
cat gpurun_out/bench_reference.txt
