#!/bin/bash
# ncu capture of the dense (PositiveUpwind) kernel, config-5 input at 48^3, coloured scatter (8 launches per pass, 4 passes)
T=${1:-r3a}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fv1_dense -s 25 -c 1 -o gpurun_out/${T}_dense48 python tools/prof_dense.py 48 colored > gpurun_out/${T}_dense_ncu.log 2>&1
timeout 300 python tools/prof_dense.py 96 colored > gpurun_out/${T}_dense96.txt 2>&1
timeout 300 python tools/prof_dense.py 96 atomic >> gpurun_out/${T}_dense96.txt 2>&1
echo done
