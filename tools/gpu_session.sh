#!/bin/bash
# One GPU session = one `gpurun --timeout S -- 'bash tools/gpu_session.sh'` from /root/repo.  This is the round-end check of the
# shipped defaults (what the driver runs: pytest -m gpu, smoke, the default bench); per-experiment sessions replace the body.
O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu -rs > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open('$O/bench.json'))
a = d['also']['4k']
print('1080p value %.0f e2e %.0f roofline %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac']))
print('4k    value %.0f e2e %.0f roofline %.3f' % (a['value'], a['e2e']['value'], a['roofline']['frac']))
print('parity', d['parity']['psnr_db'], d['parity']['max_abs_diff'], a['parity']['psnr_db'], a['parity']['max_abs_diff'])
print('e2e_process', d.get('e2e_process'))
PY
