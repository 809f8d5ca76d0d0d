#!/bin/bash
# Round-2 twenty-eighth GPU session: compute-sanitizer and the randomised parity soak on the paint build.
set -u
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done" $O/sanitize_$tool.log | tail -n 2
done
timeout 400 python tools/soak_fuzz.py 150 21 > $O/soak_21.log 2>&1; tail -n 3 $O/soak_21.log
timeout 300 python tools/soak_fuzz.py 100 22 tall > $O/soak_22.log 2>&1; tail -n 3 $O/soak_22.log
