#!/bin/bash
# Round-2 twenty-sixth GPU session: GPU suite, smoke and the bench line on the paint build.
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 900 python bench.py > $O/bench_1080p.json 2> $O/bench_1080p.err; tail -c 600 $O/bench_1080p.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_1080p.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "value_full_evaluation", "gpu_launches")}, d["parity_checked"]["differing_pixels"], d["e2e"]["value"])
print(json.dumps(d["kernels"], indent=None)[:1500])
c = d["configs"]["4k"]; print("4k", c["value"], c["ms_per_step"], c["parity_checked"]["differing_pixels"], c["e2e"]["value"], c["single_frame_us"])
print("single", d["single_frame_us"])
PY
