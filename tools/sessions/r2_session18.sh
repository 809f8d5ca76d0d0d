#!/bin/bash
# Round-2 eighteenth GPU session: paint kernel variants -- table entries two trips ahead (1), gathers pipelined across trips (4, 5).
set -u
O=gpurun_out
mkdir -p $O
for v in 1 4 5; do
  FSB_PAINT_VARIANT=$v timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint_ or batch_paths or config_2 or cfg2 or 1080" > $O/pytest_v$v.log 2>&1; tail -n 3 $O/pytest_v$v.log
done
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for v in 0 1 4 5; do
  FSB_PAINT_VARIANT=$v run 1080p 512
  FSB_PAINT_VARIANT=$v run 4k 128
done
python tools/show_variants.py $O/variants.jsonl
FSB_PAINT_VARIANT=4 FSB_PAINT_SEG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_paint --launch-skip 2 -c 1 -f -o $O/r2k_paint_v4_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_paint.log 2>&1
tail -n 2 $O/ncu_paint.log
