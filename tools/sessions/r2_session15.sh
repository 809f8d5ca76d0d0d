#!/bin/bash
# Round-2 fifteenth GPU session: paint kernel v2 (row masks, predicated PTX row body) -- parity, A/B, lane utilisation, ncu.
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
for seg in 9 17; do FSB_PAINT_SEG=$seg run 4k 128; done
for p in 128 256; do
  FSB_PAINT=0 run 1080p $p
  run 1080p $p
done
python tools/show_variants.py $O/variants.jsonl
FSB_PAINT_SEG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_paint --launch-skip 2 -c 1 -f -o $O/r2i_paint_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_paint.log 2>&1
tail -n 2 $O/ncu_paint.log
