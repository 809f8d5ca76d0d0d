#!/bin/bash
# Round-2 nineteenth GPU session: paint kernel -- candidate words prefetched into L2 (variants 1, 5); launch groups small enough
# for the candidate lists to stay in L2.
set -u
O=gpurun_out
mkdir -p $O
for v in 1 5; do
  FSB_PAINT_VARIANT=$v timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint_ or batch_paths" > $O/pytest_v$v.log 2>&1; tail -n 3 $O/pytest_v$v.log
done
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for v in 0 1 5; do
  FSB_PAINT_VARIANT=$v run 1080p 512
  FSB_PAINT_VARIANT=$v run 4k 128
done
for g in 64 128 256; do
  FSB_GROUP_POSES=$g run 1080p 512
  FSB_GROUP_POSES=$g FSB_PAINT=0 run 1080p 512
done
FSB_GROUP_POSES=64 run 4k 128
python tools/show_variants.py $O/variants.jsonl
