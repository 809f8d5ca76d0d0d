#!/bin/bash
# Round-2 forty-second GPU session: single-frame march with the two gather sets swapping roles (no copy at the end of a round).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for i in 1 2; do for wl in 1080p cfg1; do run $wl 1 0 20; done; done
run 1080p 2 0 10
run 1080p 8 0 10
python tools/show_variants.py $O/variants.jsonl
