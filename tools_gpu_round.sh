#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/all_gpu_tests.txt
tail -4 gpurun_out/all_gpu_tests.txt
(timeout 300 ncu --set full --clock-control none -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_1080p_plain python tools/profile_tc.py 1080p 0 > gpurun_out/ncu_tc_stdout.txt 2>&1)
(timeout 300 ncu --set full --clock-control none -k regex:tc_conv3x3 -s 2 -c 1 -o gpurun_out/prof_tc_conv64_4k_plain python tools/profile_tc.py 4k 0 >> gpurun_out/ncu_tc_stdout.txt 2>&1)
(timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1) > gpurun_out/bench_1080p_default.txt
cat gpurun_out/bench_1080p_default.txt
python - <<'PY'
import numpy as np, __graft_entry__ as g
pkg = g.load_package()
for split in (0, 1):
    t = pkg.debug_conv_timeline(64, 64, 272, 480, bool(split))
    r = t[0].astype(np.int64); b = r[0]
    f = lambda a: [int(x - b) if x else -1 for x in a]
    print("split", split, "mma full seen", f(r[16:24]), "committed", f(r[32:40]), "epi", f(r[44:48]))
PY
