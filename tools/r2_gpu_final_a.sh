#!/bin/bash
# round-2 evidence run: j0 vs on-chip recompute (general rows kernel), tile kernel at the contract size, all five configurations
T=${1:-r2r}
mkdir -p gpurun_out
NSB_NOSPLIT=1 timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_n128_nosplit.json 2> gpurun_out/${T}.err
NSB_NOSPLIT=1 timeout 600 ncu --set full --clock-control none -k regex:fv1_ -s 2 -c 2 -o gpurun_out/${T}_nosplit_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_nosplit.log 2>&1
NSB_TILE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_tile.json 2>> gpurun_out/${T}.err
NSB_TILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_ -s 1 -c 1 -o gpurun_out/${T}_tile_n184 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_tile.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_ -s 2 -c 2 -o gpurun_out/${T}_split_n184 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_split.log 2>&1
timeout 900 python tools/config_bench.py > gpurun_out/${T}_configs.txt 2>&1
echo done
