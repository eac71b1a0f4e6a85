#!/bin/bash
# round 2, GPU call A: parity suite, fused vs split bench on the same box, ncu of the fused kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2a_gputest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err
NSB_NOFUSED=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_split.json 2> gpurun_out/r2a_bench_split.err
NSB_RAYFAST=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_fused_norayfast.json 2> gpurun_out/r2a_bench_norayfast.err
timeout 900 python tools/quick_bench.py 96 > gpurun_out/r2a_quick96.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fv1_fused -s 1 -c 1 -o gpurun_out/r2a_fused_n128 python bench.py --cells 128 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2a_ncu.log 2>&1
echo done
