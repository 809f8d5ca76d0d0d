#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 900 python bench.py                                   > $O/bench_1080p.json 2> $O/bench_1080p.err
timeout 900 python bench.py --workload cfg1 --steps 30        > $O/bench_cfg1.json  2> $O/bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --workload 8k-colsplit --steps 10 > $O/bench_colsplit_n1.json 2> $O/bench_colsplit_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
# full captures of the two kernels of the batch path at the batch sizes bench.py uses (whole columns per paint warp at 1080p)
for k in marchc paint; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_$k --launch-skip 1 -c 1 -f -o $O/r2_${k}_1080p_b512 \
      python tools/prof_batch.py 1080p 512 1 > $O/ncu_${k}_1080p.log 2>&1
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_$k --launch-skip 1 -c 1 -f -o $O/r2_${k}_4k_b128 \
      python tools/prof_batch.py 4k 128 1 > $O/ncu_${k}_4k.log 2>&1
done
# the two launches the paint kernel replaced (A/B record)
for k in colour expand; do
  FSB_PAINT=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_$k --launch-skip 1 -c 1 -f -o $O/r2_${k}_1080p_b512 \
      python tools/prof_batch.py 1080p 512 1 > $O/ncu_${k}_1080p.log 2>&1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_march4 -c 1 -f -o $O/r2_march4_1080p_single \
    python tools/prof_batch.py 1080p 1 1 > $O/ncu_march4.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_march_kernel -c 1 -f -o $O/r2_marchz_4k_single \
    python tools/prof_batch.py 4k 1 1 > $O/ncu_marchz.log 2>&1
tail -n 1 $O/ncu_*.log | grep -v "^$" | tail -n 20
python - <<'PY'
import json
for f in ("bench_1080p", "bench_cfg1", "bench_reference", "bench_colsplit_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "value_full_evaluation")}, (d.get("parity_checked") or {}).get("differing_pixels"), (d.get("e2e") or {}).get("value"))
        if d.get("configs"):
            c = d["configs"]["4k"]; print("   4k", c["value"], c["ms_per_step"], c["parity_checked"]["differing_pixels"], c["e2e"]["value"], c["single_frame_us"])
        if d.get("single_frame_us"): print("   single", d["single_frame_us"])
    except Exception as e:
        print(f, "FAILED", e)
PY
