#!/bin/bash
# Round-2 thirty-seventh GPU session: expand with the band's records staged in shared memory (fsb_expand4s_kernel).
set -u
O=gpurun_out
mkdir -p $O
FSB_EXPAND_STAGE=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_stage1.log 2>&1; tail -n 2 $O/pytest_stage1.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in 1080p cfg1 4k; do
  for st in 0 1; do FSB_EXPAND_STAGE=$st run $wl 1 0 20; done
done
for p in 2 8 16 32; do
  for st in 0 1; do FSB_EXPAND_STAGE=$st run 1080p $p 0 10; done
done
python tools/show_variants.py $O/variants.jsonl
