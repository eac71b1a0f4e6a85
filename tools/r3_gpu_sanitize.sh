#!/bin/bash
# compute-sanitizer over a handful of small assemblies through every kernel family (default paths and the opt-in ones)
T=${1:-r3s}
mkdir -p gpurun_out
timeout 120 python tools/sanitize_cases.py > gpurun_out/${T}_plain.txt 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/${T}_$tool.txt 2>&1
  echo "== $tool default: $(grep -c '^ok' gpurun_out/${T}_$tool.txt) cases ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_$tool.txt)"
done
for v in NSB_NOFUSED=1 NSB_TILE=1 NSB_SPLIT_OWNER=0; do
  for tool in memcheck racecheck; do
    env $v timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/${T}_${tool}_$v.txt 2>&1
    echo "== $tool $v: $(grep -c '^ok' gpurun_out/${T}_${tool}_$v.txt) cases ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_${tool}_$v.txt)"
  done
done
echo done
