#!/bin/bash
T=${1:-r2n}
mkdir -p gpurun_out
: > gpurun_out/${T}_sweep.txt
for mb in 8 10; do for tg in 2 4 8; do
echo "MINB=$mb TG=$tg" >> gpurun_out/${T}_sweep.txt
NSB_OWNER_MINB=$mb NSB_TICKET_GROUP=$tg timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" >> gpurun_out/${T}_sweep.txt
done; done
echo "OWNER=0" >> gpurun_out/${T}_sweep.txt
NSB_SPLIT_OWNER=0 timeout 300 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" >> gpurun_out/${T}_sweep.txt
echo "184 default" >> gpurun_out/${T}_sweep.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" >> gpurun_out/${T}_sweep.txt
echo "184 zorder" >> gpurun_out/${T}_sweep.txt
NSB_ZORDER=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>> gpurun_out/${T}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" >> gpurun_out/${T}_sweep.txt
echo done
