#!/bin/bash
# Round 2, GPU session 5 (2 GPUs): the reference CLI in its in-process multi-GPU mode, torchrun bench weak / strong at N = 2.
O=gpurun_out/r2_s5
mkdir -p $O
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/summary.txt; }
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $O/gpus.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_cli_dropin_gpu.py tests/test_dist_cpu.py -q -p no:cacheprovider > $O/pytest_cli.log 2>&1
stamp "pytest CLI drop-in (incl. -g 0,1) + dist: rc=$? $(tail -1 $O/pytest_cli.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > $O/bench_n2_weak.json 2> $O/bench_n2_weak.err
stamp "bench N=2 weak rc=$? $(cut -c1-160 $O/bench_n2_weak.json)"
timeout 600 $TR bench.py --gpus 2 --scaling strong --workload 4k --only --no-cpu-baseline > $O/bench_n2_strong_4k.json 2> $O/bench_n2_strong_4k.err
stamp "bench N=2 strong 4K rc=$? $(cut -c1-160 $O/bench_n2_strong_4k.json)"
timeout 300 python bench.py --only --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
stamp "bench N=1 (--only, with the process leg) rc=$? $(cut -c1-160 $O/bench_n1.json)"
timeout 300 python bench.py --only --no-cpu-baseline --no-process-leg --scaling strong --workload 4k > $O/bench_n1_strong_4k.json 2> $O/bench_n1_strong_4k.err
stamp "bench N=1 strong 4K rc=$? $(cut -c1-160 $O/bench_n1_strong_4k.json)"
cat $O/summary.txt
