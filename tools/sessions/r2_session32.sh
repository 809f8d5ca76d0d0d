#!/bin/bash
# Round-2 thirty-second GPU session: the local occlusion bound per quarter of a chunk (8 steps) in the column-parallel march.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in "1080p 512" "4k 128" "cfg1 512"; do
  run $wl
  FSB_MARCHC_VARIANT=1 run $wl
  run $wl 4
done
python tools/show_variants.py $O/variants.jsonl
timeout 300 python tools/soak_fuzz.py 60 31 > $O/soak_31.log 2>&1; tail -n 2 $O/soak_31.log
