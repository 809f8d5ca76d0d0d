#!/bin/bash
# Round-2 sixth GPU session: full suite after the shim / ADVICE changes, ncu captures for profiles/r2_ncu_rates.json, launch list.
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 6 $O/pytest.log
for k in marchc colour expand; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_$k -c 1 -f -o $O/r2e_${k}_1080p_b128 \
      python tools/prof_batch.py 1080p 128 1 > $O/ncu_${k}_1080p.log 2>&1
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_$k -c 1 -f -o $O/r2e_${k}_4k_b64 \
      python tools/prof_batch.py 4k 64 1 > $O/ncu_${k}_4k.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_march_kernel -c 1 -f -o $O/r2e_marchz_1080p_single \
    python tools/prof_batch.py 1080p 1 1 > $O/ncu_marchz_single.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
tail -n 1 $O/ncu_*.log | tail -n 30
timeout 900 python bench.py > $O/bench_1080p.json 2> $O/bench_1080p.err; tail -c 300 $O/bench_1080p.err
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
run 1080p 512
run 4k 64
run 1080p 1
python tools/show_variants.py $O/variants.jsonl
ls -la $O | head -40
