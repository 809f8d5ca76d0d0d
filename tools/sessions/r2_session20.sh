#!/bin/bash
# Round-2 twentieth GPU session: paint kernel -- L2 prefetch distance sweep, occupancy / pipelined variants, ncu of the default.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
FSB_PAINT_VARIANT=4 timeout 600 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "paint_ or batch_paths" > $O/pytest_v4.log 2>&1; tail -n 2 $O/pytest_v4.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for pf in 0 3 5 8 12; do
  FSB_PAINT_PF=$pf run 1080p 512
  FSB_PAINT_PF=$pf run 4k 128
done
for v in 2 4; do
  for pf in 5 10; do
    FSB_PAINT_VARIANT=$v FSB_PAINT_PF=$pf run 1080p 512
    FSB_PAINT_VARIANT=$v FSB_PAINT_PF=$pf run 4k 128
  done
done
run cfg1 512
python tools/show_variants.py $O/variants.jsonl
FSB_PAINT_SEG=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_paint --launch-skip 2 -c 1 -f -o $O/r2l_paint_pf5_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_paint.log 2>&1
tail -n 2 $O/ncu_paint.log
