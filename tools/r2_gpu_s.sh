#!/bin/bash
T=${1:-r2s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_gputest.txt
timeout 900 python tools/config_bench.py > gpurun_out/${T}_configs.txt 2>&1
timeout 300 python tools/qb_gather.py 96 > gpurun_out/${T}_qb96.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}.err
echo done
