#!/bin/bash
# Round-2 twelfth GPU session: depth-parallel cluster march (fsb_march_split.cu) -- parity, then timing against the marches it replaces.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_render_gpu.py -k "single_frame_march_variants or tests_variant_golden" -x -q > $O/pytest_split.log 2>&1; tail -n 15 $O/pytest_split.log
grep -q "failed\|error\|Error" $O/pytest_split.log && exit 1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 6 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in cfg1 1080p 4k; do
  FSB_SPLIT=0 run $wl 1
  run $wl 1
  FSB_SPLIT_WARPS=32 run $wl 1
  FSB_SPLIT_WARPS=64 run $wl 1
  for sl in 4 16 32; do FSB_COLOUR_SLICE=$sl run $wl 1; done
done
for p in 2 4 8 16 32 64; do
  FSB_SPLIT=0 run 1080p $p
  run 1080p $p
  FSB_SPLIT_WARPS=32 FSB_SPLIT_MAX_GROUPS=100000 run 1080p $p
done
# local occlusion bound (pyramid of height maxima) on the batch path
for wl in "1080p 512" "4k 128" "cfg1 512"; do
  FSB_LOCAL_CULL=0 run $wl
  run $wl
  FSB_MARCHC_VARIANT=1 run $wl
  FSB_MARCHC_VARIANT=5 run $wl
done
python tools/show_variants.py $O/variants.jsonl
