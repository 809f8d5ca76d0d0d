#!/bin/bash
# Round-2 thirty-first GPU session: single frames -- depth-table entries loaded a round before the gathers that need them.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in 1080p cfg1; do run $wl 1 0 20; run $wl 1 0 20; done
FSB_FRAME_MAX_COLS=100000 run 4k 1 0 20
run 4k 1 0 20
run 1080p 2 0 20
python tools/show_variants.py $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_march4 --launch-skip 2 -c 1 -f -o $O/r2q_march4_1080p_single \
    python tools/prof_batch.py 1080p 1 3 > $O/ncu_march4.log 2>&1
tail -n 1 $O/ncu_march4.log
