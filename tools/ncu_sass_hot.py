#!/usr/bin/env python3
"""Per-SASS-instruction view of an ncu report's source page: executed counts + stall samples.
usage: ncu_sass_hot.py report.ncu-rep [min_exec]  -> prints the instruction stream with counts."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
tot_exec = sum(int(r[ci["Instructions Executed"]]) for r in rows[2:] if len(r) > 5)
tot_samp = sum(int(r[ci["# Samples"]]) for r in rows[2:] if len(r) > 5)
print("total executed %d, samples %d" % (tot_exec, tot_samp))
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
for n, r in enumerate(rows[2:]):
    if len(r) <= 5: continue
    ex = int(r[ci["Instructions Executed"]]); sm = int(r[ci["# Samples"]])
    stalls = {h[6:]: int(r[ci[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[ci[h]] not in ("", "0")}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    if mode == "all" or sm * 200 > tot_samp:
        print("%4d %9d %5.1f%% s=%6d %5.1f%%  %-60s %s" % (n, ex, 100.0 * ex / tot_exec, sm, 100.0 * sm / tot_samp, r[ci["Source"]].strip()[:60], top))
