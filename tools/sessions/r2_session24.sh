#!/bin/bash
# Round-2 twenty-fourth GPU session: march with per-warp table staging (no CTA barrier); paint at 8 CTAs per SM; ncu of the march.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
for wl in "1080p 512" "4k 128" "cfg1 512"; do
  FSB_MARCHC_PW=0 run $wl
  run $wl
  FSB_PAINT_VARIANT=8 run $wl
  FSB_MARCHC_PW=0 run $wl 4
  run $wl 4
done
python tools/show_variants.py $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_marchc --launch-skip 2 -c 1 -f -o $O/r2p_marchc_pw_1080p_b256 \
    python tools/prof_batch.py 1080p 256 > $O/ncu_marchc.log 2>&1
tail -n 2 $O/ncu_marchc.log
