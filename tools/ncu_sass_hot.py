#!/usr/bin/env python3
"""Per-SASS-instruction view of an ncu report's source page: executed counts + stall samples.
usage: ncu_sass_hot.py report.ncu-rep [all|hot] [kernel-substring]"""
import csv, subprocess, sys
rep = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "all"
sub = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split into per-kernel sections
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for sec in sections:
    if sub and sub not in sec["name"]:
        continue
    hdr = sec["rows"][0]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in sec["rows"][1:] if len(r) > 5]
    tot_exec = sum(int(r[ci["Instructions Executed"]]) for r in body)
    tot_samp = sum(int(r[ci["# Samples"]]) for r in body)
    print("== %s: total executed %d, samples %d" % (sec["name"], tot_exec, tot_samp))
    for n, r in enumerate(body):
        ex = int(r[ci["Instructions Executed"]]); sm = int(r[ci["# Samples"]])
        stalls = {h[6:]: int(r[ci[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[ci[h]] not in ("", "0")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
        if mode == "all" or sm * 200 > tot_samp:
            print("%4d %9d %5.1f%% s=%6d %5.1f%%  %-60s %s" % (n, ex, 100.0 * ex / max(tot_exec, 1), sm, 100.0 * sm / max(tot_samp, 1), r[ci["Source"]].strip()[:60], top))
