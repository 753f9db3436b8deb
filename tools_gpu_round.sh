#!/bin/bash
O=gpurun_out/r2_s26
mkdir -p $O
timeout 150 python bench.py --workload 4k --only --no-cpu-baseline --no-process-leg --steps 6 > $O/bench_distinct_4k.json 2> $O/bench_distinct_4k.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('$O/bench_distinct_4k.json'))
print('4k value %.0f e2e %.0f h2d %d d2h %d each %s' % (d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['e2e']['ms_each_step_this_rank']))
PY
tail -3 $O/bench_distinct_4k.err
