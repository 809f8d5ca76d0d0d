#!/bin/bash
# Round-2 twenty-fifth GPU session: pipelined-gather variant with the ordered loads; bands per paint warp on medium batches.
set -u
O=gpurun_out
mkdir -p $O
FSB_PAINT_VARIANT=4 timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint_ or batch_paths" > $O/pytest_v4.log 2>&1; tail -n 2 $O/pytest_v4.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in "1080p 512" "4k 128" "cfg1 512"; do
  run $wl
  FSB_PAINT_VARIANT=4 run $wl
done
for seg in 0 12 17 23; do FSB_PAINT_SEG=$seg run 4k 128; done
for seg in 0 5 9 17; do FSB_PAINT_SEG=$seg run 1080p 128; done
for seg in 0 9 17; do FSB_PAINT_SEG=$seg run 1080p 256; done
for seg in 0 6 12; do FSB_PAINT_SEG=$seg run cfg1 512; done
FSB_PAINT=0 run 1080p 128
FSB_PAINT=0 run 1080p 256
python tools/show_variants.py $O/variants.jsonl
