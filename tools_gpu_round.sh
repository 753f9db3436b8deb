#!/bin/bash
O=gpurun_out/r2_s23
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 420 $CS --tool memcheck --error-exitcode 9 --log-file $O/memcheck_hbm.txt python -m pytest tests/test_hbm_kernels_gpu.py -x -q -m gpu > $O/memcheck_hbm_pytest.txt 2>&1; echo "hbm unit tests under memcheck: rc=$?"; tail -2 $O/memcheck_hbm_pytest.txt; tail -3 $O/memcheck_hbm.txt
timeout 420 $CS --tool memcheck --error-exitcode 9 --log-file $O/memcheck_smoke.txt python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/memcheck_smoke_out.txt 2>&1; echo "smoke (fused v4.6 path, tcgen05 convs) under memcheck: rc=$?"; tail -2 $O/memcheck_smoke_out.txt; tail -3 $O/memcheck_smoke.txt
timeout 420 $CS --tool memcheck --error-exitcode 9 --log-file $O/memcheck_tta.txt python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "test_tta_modes" > $O/memcheck_tta_pytest.txt 2>&1; echo "TTA parity (generic executor, lanes) under memcheck: rc=$?"; tail -2 $O/memcheck_tta_pytest.txt; tail -3 $O/memcheck_tta.txt
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
