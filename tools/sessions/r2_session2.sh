#!/bin/bash
# Round-2 second GPU session: smem-staged march + rewritten colour pass + overlap mode.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 5 $O/pytest.log
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
FSB_MARCH_Z=1 run 1080p 256
for v in 0 1 2 3 6; do FSB_MARCHC_VARIANT=$v run 1080p 256; done
run 1080p 256 4
for ov in 16 32 64; do FSB_OVERLAP=$ov run 1080p 256; FSB_OVERLAP=$ov FSB_MARCHC_VARIANT=3 run 1080p 256; done
FSB_OVERLAP=32 run 1080p 512
run 1080p 512
run 4k 64
FSB_MARCHC_VARIANT=3 run 4k 64
FSB_OVERLAP=8 run 4k 64
FSB_OVERLAP=16 run 4k 64
FSB_MARCH_Z=1 run 1080p 1
for s in 8 16 32; do FSB_SEGMENTS=$s run 1080p 1; FSB_SEGMENTS=$s run 4k 1; done
FSB_SEGMENTS=16 FSB_MARCHC_VARIANT=6 run 1080p 1
FSB_SEGMENTS=16 FSB_MARCHC_VARIANT=1 run 1080p 1
FSB_SEGMENTS=8 FSB_MARCHC_VARIANT=1 run 4k 1
cat $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_marchc -c 1 -f -o $O/r2b_marchc_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_colour -c 1 -f -o $O/r2b_colour_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_ -c 5 -f -o $O/r2b_single_1080p \
    python tools/prof_batch.py 1080p 1 1 > $O/ncu3.log 2>&1
tail -n 2 $O/ncu1.log $O/ncu2.log $O/ncu3.log
