#!/bin/bash
# Round-2 forty-fourth GPU session: programmatic dependent launch for small batches on the lanes-over-depth path.
set -u
O=gpurun_out
mkdir -p $O
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for p in 2 4 8 16 32; do
  run 1080p $p 0 10
  FSB_PDL_BATCH=1 run 1080p $p 0 10
done
for p in 2 8; do
  run 4k $p 0 10
  FSB_PDL_BATCH=1 run 4k $p 0 10
done
python tools/show_variants.py $O/variants.jsonl
FSB_PDL_BATCH=1 timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "batch or variants or fuzz" > $O/pytest_pdlb.log 2>&1; tail -n 2 $O/pytest_pdlb.log
