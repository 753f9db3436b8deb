mkdir -p gpurun_out/s12
timeout 600 python -m pytest tests/test_tc_conv_gpu.py -x -q > gpurun_out/s12/tc_tests.log 2>&1; tail -3 gpurun_out/s12/tc_tests.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x > gpurun_out/s12/parity.log 2>&1; tail -4 gpurun_out/s12/parity.log
RIFE_B200_KTIME=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --lanes 1 > gpurun_out/s12/bench_l1.json 2> gpurun_out/s12/ktime.txt; grep ktime gpurun_out/s12/ktime.txt | grep -v "conv[01]\|b[01] "
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s12/bench.json 2> gpurun_out/s12/bench.err; cat gpurun_out/s12/bench.json | cut -c1-200; grep -o '"e2e": {[^}]*}' gpurun_out/s12/bench.json; grep -o '"clocks": {[^}]*}' gpurun_out/s12/bench.json
