#!/bin/bash
# Round-2 first GPU session: parity of the column-parallel march, instruction-rate microbenchmarks, variant sweep, ncu.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 15 $O/pytest.log
timeout 120 tools/scratch/bin/tex_rate > $O/tex_rate.txt 2>&1; tail -n 8 $O/tex_rate.txt
: > $O/variants.jsonl
FSB_MARCH_Z=1 timeout 300 python tools/r2_time.py 1080p 256 >> $O/variants.jsonl 2>> $O/variants.err
for v in 0 1 2 3 4 6; do
  FSB_MARCHC_VARIANT=$v timeout 300 python tools/r2_time.py 1080p 256 >> $O/variants.jsonl 2>> $O/variants.err
done
timeout 300 python tools/r2_time.py 1080p 256 4 >> $O/variants.jsonl 2>> $O/variants.err
FSB_MARCH_Z=1 timeout 300 python tools/r2_time.py 4k 64 >> $O/variants.jsonl 2>> $O/variants.err
timeout 300 python tools/r2_time.py 4k 64 >> $O/variants.jsonl 2>> $O/variants.err
FSB_MARCHC_VARIANT=2 timeout 300 python tools/r2_time.py 4k 64 >> $O/variants.jsonl 2>> $O/variants.err
# single frames: segments sweep
FSB_MARCH_Z=1 timeout 300 python tools/r2_time.py 1080p 1 >> $O/variants.jsonl 2>> $O/variants.err
FSB_MARCH_Z=1 timeout 300 python tools/r2_time.py 4k 1 >> $O/variants.jsonl 2>> $O/variants.err
for s in 1 4 8 16 32; do
  FSB_SEGMENTS=$s timeout 300 python tools/r2_time.py 1080p 1 >> $O/variants.jsonl 2>> $O/variants.err
  FSB_SEGMENTS=$s timeout 300 python tools/r2_time.py 4k 1 >> $O/variants.jsonl 2>> $O/variants.err
done
FSB_SEGMENTS=16 FSB_MARCHC_VARIANT=5 timeout 300 python tools/r2_time.py 1080p 1 >> $O/variants.jsonl 2>> $O/variants.err
FSB_SEGMENTS=16 FSB_MARCHC_VARIANT=5 timeout 300 python tools/r2_time.py 4k 1 >> $O/variants.jsonl 2>> $O/variants.err
cat $O/variants.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_marchc -c 1 -f -o $O/r2_marchc_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_colour -c 1 -f -o $O/r2_colour_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu2.log 2>&1
tail -n 2 $O/ncu1.log $O/ncu2.log
