#!/bin/bash
# tile kernel bring-up: parity suite, contract bench, bench at 128^3 (tile vs split), ncu of the tile kernel
T=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_gputest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_n128.json 2>> gpurun_out/${T}_bench.err
NSB_TILE=0 timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_n128_split.json 2>> gpurun_out/${T}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_ -s 2 -c 1 -o gpurun_out/${T}_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu.log 2>&1
echo done
