#!/bin/bash
T=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_gputest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python tools/quick_bench.py 96 > gpurun_out/${T}_quick96.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_ -s 2 -c 2 -o gpurun_out/${T}_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu.log 2>&1
echo done
