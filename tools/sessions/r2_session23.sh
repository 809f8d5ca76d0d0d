#!/bin/bash
# Round-2 twenty-third GPU session: paint kernel -- gathers issued before the next-trip loads and the bookkeeping.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint_ or batch_paths" > $O/pytest_p.log 2>&1; tail -n 2 $O/pytest_p.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for v in 0 2; do
  FSB_PAINT_VARIANT=$v run 1080p 512
  FSB_PAINT_VARIANT=$v run 4k 128
  FSB_PAINT_VARIANT=$v run cfg1 512
done
python tools/show_variants.py $O/variants.jsonl
FSB_PAINT_SEG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_paint --launch-skip 2 -c 1 -f -o $O/r2o_paint_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_paint.log 2>&1
tail -n 2 $O/ncu_paint.log
