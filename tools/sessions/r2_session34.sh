#!/bin/bash
# Round-2 thirty-fourth GPU session: GPU suite with the batch path from 16 warps of 32 columns per SM; timings around the threshold.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for p in 36 40 44; do run 1080p $p 0 10; done
for p in 18 20 24; do run 4k $p 0 10; done
FSB_COLS_MIN_WARPS=100000000 run 1080p 40 0 10
FSB_COLS_MIN_WARPS=100000000 run 4k 20 0 10
python tools/show_variants.py $O/variants.jsonl
