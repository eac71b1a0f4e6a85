#!/bin/bash
# final evidence run of the round on one B200: GPU test suite, bench line, reference arm, ncu launch list of the bench command, one
# --set full capture of the two kernels of a pass at the contract size, the five configurations, cavity vs the reference's Ghia tables
T=${1:-r3x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_gputest.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:fv1_ -s 3 -c 2 -o gpurun_out/${T}_split_n184 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/${T}_ncu_split.log 2>&1
timeout 900 python tools/config_bench.py > gpurun_out/${T}_configs.txt 2>&1
NSB_CAVITY_ELEM=tri NSB_CAVITY_JITTER=0.2 timeout 300 python tools/cavity_ghia.py 100 32 64 > gpurun_out/${T}_cavity_tri.txt 2>&1
NSB_CAVITY_ELEM=quad NSB_CAVITY_JITTER=0.2 timeout 300 python tools/cavity_ghia.py 100 64 >> gpurun_out/${T}_cavity_tri.txt 2>&1
echo done
