#!/bin/bash
# Round-2 eighth GPU session: single-frame march with four warps per column.
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 6 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in cfg1 1080p 4k; do
  for p in 1 2 4 8; do
    FSB_FRAME_MAX_COLS=0 run $wl $p
    FSB_FRAME_MAX_COLS=100000000 run $wl $p
  done
done
run 1080p 512
run 4k 64
python tools/show_variants.py $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_march4 -c 1 -f -o $O/r2f_march4_1080p_single \
    python tools/prof_batch.py 1080p 1 1 > $O/ncu_march4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_expand4 -c 1 -f -o $O/r2f_expand4_1080p_single \
    python tools/prof_batch.py 1080p 1 1 > $O/ncu_expand4s.log 2>&1
tail -n 1 $O/ncu_march4.log $O/ncu_expand4s.log
