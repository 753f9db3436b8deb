#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_tc_conv_gpu.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/tc_conv_tests.txt
tail -2 gpurun_out/tc_conv_tests.txt
python - <<'PY'
import torch, numpy as np, __graft_entry__ as g
pkg = g.load_package()
s = torch.cuda.Stream()
for (c, h, w, split) in [(64, 272, 480, 0), (64, 272, 480, 1), (64, 544, 960, 0), (64, 544, 960, 1), (96, 136, 240, 0), (128, 68, 120, 1), (192, 34, 60, 1)]:
    with torch.cuda.stream(s):
        pkg.bench_conv(s.cuda_stream, c, c, h, w, split, 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); pkg.bench_conv(s.cuda_stream, c, c, h, w, split, 20); e1.record(s)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1000
    print("conv %d %dx%d split=%d: %.2f us  %.1f TFLOP/s useful" % (c, w, h, split, us, 2 * 9 * c * c * h * w / us / 1e6))
for split in (0, 1):
    t = pkg.debug_conv_timeline(64, 64, 272, 480, bool(split))
    r = t[0].astype(np.int64); b = r[0]
    f = lambda a: [int(x - b) if x else -1 for x in a]
    print("split", split, "mma full seen", f(r[16:24]), "committed", f(r[32:40]), "epi", f(r[44:48]))
PY
(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_1080p_default.txt
python -c "import json;d=json.load(open('gpurun_out/bench_1080p_default.txt'));print('1080p',round(d['value'],1),round(d['e2e']['value'],1),d['roofline']['us_per_launch'])"
(timeout 300 python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_4k_default.txt
python -c "import json;d=json.load(open('gpurun_out/bench_4k_default.txt'));print('4k',round(d['value'],1),round(d['e2e']['value'],1),d['roofline']['us_per_launch'])"
