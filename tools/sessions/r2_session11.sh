#!/bin/bash
# Round-2 eleventh GPU session: programmatic dependent launch on the single-frame path.
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 4 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in cfg1 1080p 4k; do
  FSB_PDL=0 run $wl 1
  run $wl 1
done
run 1080p 512
python tools/show_variants.py $O/variants.jsonl
timeout 300 python tools/soak_fuzz.py 60 21 > $O/soak_21.log 2>&1; tail -n 2 $O/soak_21.log
