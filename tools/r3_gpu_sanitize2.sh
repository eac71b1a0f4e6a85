#!/bin/bash
# compute-sanitizer over the GPU tests of the kernels beside the element assembly (boundary discs, Dirichlet post-pass, resident
# Jacobian / SpMV, per-ip data, turbulence / diagnostics, FVCR constraint, phased assembly)
T=${1:-r3u}
mkdir -p gpurun_out
FILES="tests/test_gpu_boundary.py tests/test_gpu_turbulence.py tests/test_constraint_fvcr.py tests/test_gpu_phased.py tests/test_ip_data.py tests/test_gpu_resident.py"
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest $FILES -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_memcheck_tests.txt 2>&1
echo "== memcheck: $(grep -E 'passed|failed' gpurun_out/${T}_memcheck_tests.txt | tail -1); $(grep -E 'ERROR SUMMARY' gpurun_out/${T}_memcheck_tests.txt)"
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_turbulence.py tests/test_constraint_fvcr.py tests/test_gpu_phased.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${T}_racecheck_tests.txt 2>&1
echo "== racecheck: $(grep -E 'passed|failed' gpurun_out/${T}_racecheck_tests.txt | tail -1); $(grep -E 'RACECHECK SUMMARY' gpurun_out/${T}_racecheck_tests.txt)"
echo done
