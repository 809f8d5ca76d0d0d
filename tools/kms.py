"""Print value + per-kernel ms of a bench.py JSON line (development helper)."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    dp = d.get("roofline", {}).get("default_path", {})
    print("value %.0f %s  ms/step %.3f  default kernels %s  nocull kernels %s" % (
        d["value"], d["unit"], d["ms_per_step"], dp.get("kernel_ms_per_step"),
        d.get("roofline_step", {}).get("kernel_ms_per_step")))
