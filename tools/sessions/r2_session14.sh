#!/bin/bash
# Round-2 fourteenth GPU session: colour pass + expand as one kernel (fsb_paint.cu) -- parity, then A/B against the two launches.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 15 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in "1080p 512" "4k 128" "cfg1 512"; do
  FSB_PAINT=0 run $wl
  run $wl
  FSB_PAINT_SEG=0 run $wl
done
for seg in 4 9 17; do FSB_PAINT_SEG=$seg run 4k 128; done
for p in 64 128 256; do
  FSB_PAINT=0 run 1080p $p
  run 1080p $p
  FSB_PAINT_SEG=0 run 1080p $p
done
python tools/show_variants.py $O/variants.jsonl
