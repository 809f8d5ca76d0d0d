#!/bin/bash
# Round-2 thirty-ninth GPU session: shared-memory carve-out of the single-frame chain.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for c in 100 75 50 35; do
  for wl in 1080p cfg1 4k; do FSB_CARVEOUT=$c run $wl 1 0 20; done
done
python tools/show_variants.py $O/variants.jsonl
