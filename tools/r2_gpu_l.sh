#!/bin/bash
T=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_gputest.txt
for mb in 12 10 16; do
NSB_OWNER_MINB=$mb timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_n128_mb$mb.json 2>> gpurun_out/${T}_bench.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench.json 2>> gpurun_out/${T}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_ -s 2 -c 2 -o gpurun_out/${T}_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu.log 2>&1
echo done
