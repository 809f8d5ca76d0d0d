#!/usr/bin/env python3
"""Per-pose rates of the render kernels (march, paint; colour and expand for the two-launch A/B) from ncu captures of tools/prof_batch.py, merged into
profiles/r2_ncu_rates.json (read by bench.py for issue_slot_frac / L2 sectors per sample / DRAM traffic).
usage: ncu_rates.py <workload> <counters.json> march=<rep> paint=<rep> [colour=<rep> expand=<rep>]"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"inst": "smsp__inst_executed.sum", "dram_r": "dram__bytes_read.sum", "dram_w": "dram__bytes_write.sum",
        "lts_tex_read": "lts__t_sectors_srcunit_tex_op_read.sum", "dur": "gpu__time_duration.sum",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def first_kernel(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    d = {"kernel": r[col["Kernel Name"]]}
    for k, name in KEYS.items():
        v = float(r[col[name]].replace(",", ""))
        d[k] = v * SCALE.get(units[col[name]], 1.0)
    return d


def main():
    wl, counters = sys.argv[1], json.load(open(sys.argv[2]))
    P = counters["poses"]
    path = os.path.join(ROOT, "profiles", "r2_ncu_rates.json")
    allr = json.load(open(path)) if os.path.exists(path) else {}
    e = {"poses_in_capture": P, "source": "ncu --set full --clock-control none, tools/prof_batch.py %s %d" % (wl, P)}
    for a in sys.argv[3:]:
        name, rep = a.split("=")
        k = first_kernel(rep)
        e[name] = {"kernel": k["kernel"], "inst_per_pose": k["inst"] / P, "dram_bytes_per_pose": (k["dram_r"] + k["dram_w"]) / P,
                   "lts_tex_read_sectors_per_pose": k["lts_tex_read"] / P, "duration_us_cold": k["dur"],
                   "issue_active_pct": k["issue"], "report": os.path.basename(rep)}
        if name == "march":
            e[name]["chunks_frac_in_capture"] = counters["chunks_frac"]
    allr[wl] = e
    json.dump(allr, open(path, "w"), indent=1)
    print(json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
