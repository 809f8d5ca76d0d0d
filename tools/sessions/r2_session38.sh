#!/bin/bash
# Round-2 thirty-eighth GPU session: single frames with one shared-memory carve-out for the whole chain; staged expand A/B.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in 1080p cfg1 4k; do
  for st in 0 1; do FSB_EXPAND_STAGE=$st run $wl 1 0 20; done
done
for p in 2 4 8; do
  for st in 0 1; do FSB_EXPAND_STAGE=$st run 1080p $p 0 10; done
done
python tools/show_variants.py $O/variants.jsonl
