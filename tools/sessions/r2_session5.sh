#!/bin/bash
# Round-2 fifth GPU session: full test suite, the bench line (1080p + 4K legs, parity checks), the CPU arm, column split at N = 1.
set -u
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 5 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
timeout 900 python bench.py > $O/bench_1080p.json 2> $O/bench_1080p.err; tail -c 600 $O/bench_1080p.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 900 python bench.py --workload 8k-colsplit --steps 10 > $O/bench_colsplit_n1.json 2> $O/bench_colsplit_n1.err; tail -c 600 $O/bench_colsplit_n1.err
: > $O/variants.jsonl
run() { timeout 300 python tools/r2_time.py "$@" >> $O/variants.jsonl 2>> $O/variants.err; }
run 1080p 256
run 1080p 512
run 4k 64
FSB_COLOUR_SLICE=64 run 4k 64
run 4k 32
FSB_MARCH_Z=1 run 4k 32
run 1080p 96
FSB_MARCH_Z=1 run 1080p 96
cat $O/variants.jsonl
python - <<'PY'
import json
for f in ("bench_1080p", "bench_reference", "bench_colsplit_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "value_full_evaluation", "parity_checked")}, d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
