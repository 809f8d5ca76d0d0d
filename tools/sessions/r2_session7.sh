#!/bin/bash
# Round-2 seventh GPU session: expand on a high-priority stream beside the next group's march (overlap experiment).
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
run 1080p 512
for ov in 32 64 128; do
  for kb in 0 40 72 100; do
    FSB_OVERLAP=$ov FSB_EXPAND_SMEM_KB=$kb run 1080p 512
  done
done
FSB_EXPAND_SMEM_KB=72 run 1080p 512
run 4k 64
FSB_OVERLAP=16 FSB_EXPAND_SMEM_KB=72 run 4k 64
python tools/show_variants.py $O/variants.jsonl
