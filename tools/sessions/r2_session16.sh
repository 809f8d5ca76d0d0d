#!/bin/bash
# Round-2 sixteenth GPU session: paint kernel with the ring aligned by hand -- parity; prefetch depth / occupancy variants.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 15 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for v in 0 1 2 3; do
  FSB_PAINT_VARIANT=$v run 1080p 512
  FSB_PAINT_VARIANT=$v run 4k 128
  FSB_PAINT_VARIANT=$v run cfg1 512
done
python tools/show_variants.py $O/variants.jsonl
