#!/bin/bash
# Round-2 thirtieth GPU session: single frames -- six / eight warps per column.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in 1080p cfg1; do
  for wv in 4 6 8; do FSB_FRAME_WARPS=$wv run $wl 1 0 20; done
done
for wv in 6 8; do FSB_FRAME_MAX_COLS=100000 FSB_FRAME_WARPS=$wv run 4k 1 0 20; done
python tools/show_variants.py $O/variants.jsonl
FSB_FRAME_WARPS=8 timeout 300 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "single or golden or cfg or config" > $O/pytest_w8.log 2>&1; tail -n 2 $O/pytest_w8.log
