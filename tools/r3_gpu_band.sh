#!/bin/bash
# node traversal order of the rows kernel at the contract size: natural, Z-curve, y-bands of W cells
T=${1:-r3j}
mkdir -p gpurun_out
for v in "" "NSB_ZORDER=1" "NSB_NODE_BAND=8" "NSB_NODE_BAND=16" "NSB_NODE_BAND=32" "NSB_NODE_BAND=64"; do
  echo "== ${v:-natural}" >> gpurun_out/${T}_band.txt
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_band.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/${T}_band.txt
done
echo done
