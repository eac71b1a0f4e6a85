#!/bin/bash
# one ncu --set full capture per BASELINE configuration other than config 3 (the kernels of one assembly pass after the warm-up passes)
T=${1:-r3y}
mkdir -p gpurun_out
NSB_CONFIGS=1 timeout 600 ncu --set full --clock-control none -k regex:"fv1_flux_kernel|fv1_rows_kernel" -s 6 -c 2 -o gpurun_out/${T}_config1 python tools/config_bench.py > gpurun_out/${T}_config1.log 2>&1
NSB_CONFIGS=2 timeout 600 ncu --set full --clock-control none -k regex:fv1_fused_kernel -s 3 -c 1 -o gpurun_out/${T}_config2 python tools/config_bench.py > gpurun_out/${T}_config2.log 2>&1
NSB_CONFIGS=4 timeout 600 ncu --set full --clock-control none -k regex:fvcr_elem_kernel -s 3 -c 1 -o gpurun_out/${T}_config4 python tools/config_bench.py > gpurun_out/${T}_config4.log 2>&1
NSB_CONFIGS=5 timeout 600 ncu --set full --clock-control none -k regex:fv1_dense_kernel -s 25 -c 1 -o gpurun_out/${T}_config5 python tools/config_bench.py > gpurun_out/${T}_config5.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
echo done
