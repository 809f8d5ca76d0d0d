#!/bin/bash
# Round-2 ninth GPU session: expand with TMA tile stores, A/B against the per-lane stores.
set -u
O=gpurun_out
mkdir -p $O
FSB_EXPAND_TMA=1 timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_tma.log 2>&1; tail -n 6 $O/pytest_tma.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for t in 0 1; do
  FSB_EXPAND_TMA=$t run 1080p 512
  FSB_EXPAND_TMA=$t run 4k 128
  FSB_EXPAND_TMA=$t run cfg1 512
  FSB_EXPAND_TMA=$t run 1080p 1
  FSB_EXPAND_TMA=$t run 4k 1
done
python tools/show_variants.py $O/variants.jsonl
FSB_EXPAND_TMA=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_expand4_tma -c 1 -f -o $O/r2g_expand_tma_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu_tma.log 2>&1
tail -n 1 $O/ncu_tma.log
