#!/bin/bash
O=gpurun_out/r2_s19
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "frame or cache or batch or null" > $O/pytest_frames.txt 2>&1; tail -2 $O/pytest_frames.txt
for m in 1 0 1 0; do
  RIFE_B200_D2H=$m timeout 600 python bench.py --no-cpu-baseline --no-process-leg > $O/bench_d2h$m.json 2> $O/bench_d2h$m.err
  python - <<PY
import json
d=json.load(open('$O/bench_d2h$m.json'))
a=d.get('also',{}).get('4k',{})
print('D2H=$m 1080p value %.0f e2e %.0f %s' % (d['value'], d['e2e']['value'], d['e2e']['ms_each_step_this_rank']))
print('      4k value %.0f e2e %.0f %s' % (a['value'], a['e2e']['value'], a['e2e']['ms_each_step_this_rank']))
PY
done
