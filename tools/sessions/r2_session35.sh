#!/bin/bash
# Round-2 thirty-fifth GPU session: the bench lines of the final build (1080p + 4K, cfg1, reference arm) and the launch list.
set -u
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 900 python bench.py                                   > $O/bench_1080p.json 2> $O/bench_1080p.err
timeout 900 python bench.py --workload cfg1 --steps 30        > $O/bench_cfg1.json  2> $O/bench_cfg1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/bench_under_ncu.log 2>&1
python - <<'PY'
import json
for f in ("bench_1080p", "bench_cfg1", "bench_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "value_full_evaluation")}, (d.get("parity_checked") or {}).get("differing_pixels"), (d.get("e2e") or {}).get("value"))
        if d.get("roofline"): print("   roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "launch_ms")})
        if d.get("kernels"): print("   kernels", json.dumps(d["kernels"])[:900])
        if d.get("configs"):
            c = d["configs"]["4k"]; print("   4k", c["value"], c["ms_per_step"], c["parity_checked"]["differing_pixels"], c["e2e"]["value"], c["single_frame_us"], c["roofline"]["frac"], json.dumps(c["kernels"]["paint"]))
        if d.get("single_frame_us"): print("   single", d["single_frame_us"])
    except Exception as e:
        print(f, "FAILED", e)
PY
