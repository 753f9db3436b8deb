mkdir -p gpurun_out/s19
timeout 200 python tools/mma_cost.py > gpurun_out/s19/mma_cost_commit.txt 2>&1; cat gpurun_out/s19/mma_cost_commit.txt
