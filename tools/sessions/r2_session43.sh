#!/bin/bash
# Round-2 forty-third GPU session: validation of the final build -- GPU suite, smoke, sanitizer (all three tools), parity soak.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 2 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $O/sanitize_$tool.log | tail -n 1
done
timeout 300 python tools/soak_fuzz.py 100 41 > $O/soak_41.log 2>&1; tail -n 2 $O/soak_41.log
timeout 300 python tools/soak_fuzz.py 60 42 tall > $O/soak_42.log 2>&1; tail -n 2 $O/soak_42.log
