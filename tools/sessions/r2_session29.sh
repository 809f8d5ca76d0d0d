#!/bin/bash
# Round-2 twenty-ninth GPU session: single frames -- two / three / four warps per column in fsb_march4_kernel.
set -u
O=gpurun_out
mkdir -p $O
for wv in 3 2; do
  FSB_FRAME_WARPS=$wv timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_w$wv.log 2>&1; tail -n 2 $O/pytest_w$wv.log
done
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in 1080p cfg1; do
  for wv in 4 3 2; do FSB_FRAME_WARPS=$wv run $wl 1 0 20; done
done
for wv in 4 3 2; do FSB_FRAME_MAX_COLS=100000 FSB_FRAME_WARPS=$wv run 4k 1 0 20; done
FSB_FRAME_MAX_COLS=0 run 4k 1 0 20
for wv in 4 3 2; do FSB_FRAME_MAX_COLS=100000 FSB_FRAME_WARPS=$wv run 1080p 2 0 20; done
FSB_FRAME_MAX_COLS=0 run 1080p 2 0 20
python tools/show_variants.py $O/variants.jsonl
