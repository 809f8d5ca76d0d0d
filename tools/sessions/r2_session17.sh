#!/bin/bash
# Round-2 seventeenth GPU session: paint kernel, shared-memory carve-out; occupancy seen by ncu.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for v in 0 2; do
  FSB_PAINT_VARIANT=$v run 1080p 512
  FSB_PAINT_VARIANT=$v run 4k 128
done
FSB_PAINT=0 run 1080p 512
python tools/show_variants.py $O/variants.jsonl
FSB_PAINT_SEG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_paint --launch-skip 2 -c 1 -f -o $O/r2j_paint_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_paint.log 2>&1
tail -n 2 $O/ncu_paint.log
