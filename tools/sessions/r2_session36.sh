#!/bin/bash
# Round-2 thirty-sixth GPU session: poses per step of the 4K leg.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for p in 128 192 256 384; do run 4k $p 0 5; done
FSB_PAINT_SEG=0 run 4k 256 0 5
FSB_PAINT_SEG=34 run 4k 256 0 5
python tools/show_variants.py $O/variants.jsonl
