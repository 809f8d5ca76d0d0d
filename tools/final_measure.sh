#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python bench.py                                   > $O/bench_1080p.json 2> $O/bench_1080p.err
python bench.py --workload 4k --steps 30          > $O/bench_4k.json    2> $O/bench_4k.err
python bench.py --workload cfg1 --steps 30        > $O/bench_cfg1.json  2> $O/bench_cfg1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fsb_march -c 1 -f -o $O/march_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fsb_expand -c 1 -f -o $O/expand_1080p_b128 \
    python tools/prof_batch.py 1080p 128 1 > $O/ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fsb_march -c 1 -f -o $O/march_4k_single \
    python tools/prof_batch.py 4k 1 1 > $O/ncu3.log 2>&1
for f in $O/ncu1.log $O/ncu2.log $O/ncu3.log; do tail -n 2 $f; done
