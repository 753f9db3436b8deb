#!/bin/bash
O=gpurun_out/r2_s25
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -rs -k "more_models" > $O/pytest_more_models.txt 2>&1; echo "rc=$?"; tail -5 $O/pytest_more_models.txt
