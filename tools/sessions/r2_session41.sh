#!/bin/bash
# Round-2 forty-first GPU session: single-frame march with the round's inv_z loaded a round ahead.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for i in 1 2; do for wl in 1080p cfg1; do run $wl 1 0 20; done; done
python tools/show_variants.py $O/variants.jsonl
