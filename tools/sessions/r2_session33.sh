#!/bin/bash
# Round-2 thirty-third GPU session: where the column-parallel path (march + paint) overtakes the lanes-over-depth march now.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint" > $O/pytest_p.log 2>&1; tail -n 2 $O/pytest_p.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for p in 8 16 24 32 48 64; do
  run 1080p $p 0 10
  FSB_COLS_MIN_WARPS=0 run 1080p $p 0 10
done
for p in 8 16 32; do
  run 4k $p 0 10
  FSB_COLS_MIN_WARPS=0 run 4k $p 0 10
done
python tools/show_variants.py $O/variants.jsonl
