#!/bin/bash
# Round-2 tenth GPU session: compute-sanitizer on every kernel variant, randomised parity soak on the new paths.
set -u
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/san_$tool.log 2>&1; tail -n 3 $O/san_$tool.log
  FSB_EXPAND_TMA=1 timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/san_tma_$tool.log 2>&1; tail -n 2 $O/san_tma_$tool.log
done
timeout 400 python tools/soak_fuzz.py 150 11 > $O/soak_11.log 2>&1; tail -n 3 $O/soak_11.log
timeout 400 python tools/soak_fuzz.py 100 12 tall > $O/soak_12.log 2>&1; tail -n 3 $O/soak_12.log
FSB_EXPAND_TMA=1 timeout 400 python tools/soak_fuzz.py 80 13 > $O/soak_13_tma.log 2>&1; tail -n 3 $O/soak_13_tma.log
