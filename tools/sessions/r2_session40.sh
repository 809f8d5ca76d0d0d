#!/bin/bash
# Round-2 fortieth GPU session: sanitizer on the build with the staged expand; bench lines of the final build.
set -u
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done" $O/sanitize_$tool.log | tail -n 2
done
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
timeout 900 python bench.py                            > $O/bench_1080p.json 2> $O/bench_1080p.err
timeout 900 python bench.py --workload cfg1 --steps 30 > $O/bench_cfg1.json  2> $O/bench_cfg1.err
python - <<'PY'
import json
for f in ("bench_1080p", "bench_cfg1"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("value", "ms_per_step", "value_full_evaluation")}, d["parity_checked"]["differing_pixels"], d["e2e"]["value"], d["single_frame_us"])
    if d.get("configs"):
        c = d["configs"]["4k"]; print("   4k", c["value"], c["parity_checked"]["differing_pixels"], c["e2e"]["value"], c["single_frame_us"])
PY
