#!/bin/bash
# GPU call C: parity, bench with / without the static geometry table, ncu
T=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_gputest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_fused.json 2> gpurun_out/${T}_bench.err
NSB_GEOTAB=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_fused_nogeo.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_fused_n128.json 2>> gpurun_out/${T}_bench.err
NSB_GEOTAB=0 timeout 900 python -m pytest tests/test_gpu_parity_fv1.py tests/test_gpu_workloads.py -m gpu -x -q -k "gather" 2>&1 | tail -3 > gpurun_out/${T}_gputest_nogeo.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_fused -s 1 -c 1 -o gpurun_out/${T}_fused_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu.log 2>&1
echo done
NSB_GEOTAB=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_fused -s 1 -c 1 -o gpurun_out/${T}_fused_nogeo_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_nogeo.log 2>&1
