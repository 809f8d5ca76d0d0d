#!/bin/bash
# Round-2 thirteenth GPU session: GPU suite on the local-cull build; where a single frame's time goes (ncu launch lists of
# the two single-frame paths, a full capture of the depth-parallel cluster march).
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -n 6 $O/pytest.log
for cfg in cfg2 cfg3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_${cfg}_default.csv python tools/prof_one.py $cfg live 6 > /dev/null 2>&1
  FSB_SPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_${cfg}_split.csv python tools/prof_one.py $cfg live 6 > /dev/null 2>&1
done
python - <<'PY'
import csv, glob
for f in sorted(glob.glob("gpurun_out/launches_cfg*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    print(f)
    for r in rows[-8:]:
        print("   %-60s %s %s" % (r[4][:60], r[-1], r[-2]))
PY
FSB_SPLIT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_marchs --launch-skip 3 -c 1 -f -o $O/r2h_marchs_1080p_single \
    python tools/prof_one.py cfg2 live 6 > $O/ncu_marchs.log 2>&1
tail -n 1 $O/ncu_marchs.log
FSB_SPLIT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:fsb_marchs --launch-skip 3 -c 1 -f -o $O/r2h_marchs_4k_single \
    python tools/prof_one.py cfg3 live 6 > $O/ncu_marchs4k.log 2>&1
tail -n 1 $O/ncu_marchs4k.log
