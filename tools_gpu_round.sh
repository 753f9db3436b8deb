#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_tc_conv_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/tc_conv_tests.txt
tail -4 gpurun_out/tc_conv_tests.txt
python - <<'PY' > gpurun_out/timeline.txt 2>&1
import numpy as np, __graft_entry__ as g
pkg = g.load_package()
for (c, h, w) in [(64, 272, 480), (128, 68, 120)]:
    t = pkg.debug_conv_timeline(c, c, h, w, True)
    for cta in (0, 100):
        r = t[cta].astype(np.int64)
        if r[0] == 0: continue
        b = r[0]
        f = lambda a: [int(x - b) if x else -1 for x in a]
        print("c=%d %dx%d cta %d: prod issue" % (c, w, h, cta), f(r[1:13]))
        print("   mma full seen", f(r[16:28]))
        print("   mma committed", f(r[32:44]))
        print("   epi [full,done]x4", f(r[44:52]), "end", int(r[56] - b))
PY
cat gpurun_out/timeline.txt
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q 2>&1 | tail -10) > gpurun_out/parity_tests.txt
tail -4 gpurun_out/parity_tests.txt
for lanes in 1 2; do
(timeout 300 python bench.py --steps 5 --warmup 3 --lanes $lanes --no-cpu-baseline 2>&1 | tail -2) > gpurun_out/bench_1080p_l$lanes.txt
cut -c1-420 gpurun_out/bench_1080p_l$lanes.txt; grep -o '"roofline.*us_per_launch[^,]*' gpurun_out/bench_1080p_l$lanes.txt
done
(timeout 300 python bench.py --steps 5 --warmup 3 --workload 4k --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/bench_4k.txt
cut -c1-420 gpurun_out/bench_4k.txt; grep -o '"roofline.*us_per_launch[^,]*' gpurun_out/bench_4k.txt
