#!/bin/bash
# Round-2 forty-fifth GPU session: four-way search for a paint segment's first candidate.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
run 4k 128
run 4k 128
run 1080p 128
run cfg1 512
python tools/show_variants.py $O/variants.jsonl
timeout 200 python tools/soak_fuzz.py 40 51 > $O/soak_51.log 2>&1; tail -n 1 $O/soak_51.log
