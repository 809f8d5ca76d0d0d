#!/bin/bash
# Round-2 fourth GPU session: sliced colour pass with 2-deep prefetch; crossover between the two marches.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 5 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for sl in 0 32 16 64; do FSB_COLOUR_SLICE=$sl run 1080p 256; done
for sl in 0 32; do FSB_COLOUR_SLICE=$sl run 1080p 512; done
for sl in 0 32 16; do FSB_COLOUR_SLICE=$sl run 4k 64; done
FSB_COLOUR_SLICE=32 run 4k 128
FSB_MARCH_Z=1 run 4k 128
for p in 1 2 4 8 16 32 64; do
  FSB_MARCH_Z=1 run 1080p $p
  FSB_COLS_MIN_WARPS=0 run 1080p $p
done
for p in 1 2 4 8 16; do
  FSB_MARCH_Z=1 run 4k $p
  FSB_COLS_MIN_WARPS=0 run 4k $p
done
FSB_MARCH_Z=1 run cfg1 512
run cfg1 512
cat $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_colour -c 1 -f -o $O/r2d_colour_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu2.log 2>&1
tail -n 2 $O/ncu2.log
