#!/bin/bash
T=${1:-r2q}
mkdir -p gpurun_out
: > gpurun_out/${T}_sweep.txt
run() { echo "$1" >> gpurun_out/${T}_sweep.txt; env $1 timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" >> gpurun_out/${T}_sweep.txt; }
run "NSB_FLUX_MINB=3"
run "NSB_FLUX_MINB=4"
run "NSB_FLUX_MINB=5"
run "NSB_SPLIT_WPB=1"
run "NSB_OWNER_MINB=8 NSB_TICKET_GROUP=4"
echo done
