"""Batch workload for ncu: N poses spread over the bench camera path, a few launches; afterwards the work counters of one
more step are written to gpurun_out/prof_batch_<workload>_<poses>.json (tools/ncu_rates.py pairs them with the capture).
usage: prof_batch.py <workload> <poses> [reps] [flags]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import futspace_b200 as F
import bench
key = sys.argv[1] if len(sys.argv) > 1 else "1080p"
wl = bench.WORKLOADS[key]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
ctx = F.Context(0)
col, hgt = F.terrain_fbm(m)
mp = ctx.upload_map(col, hgt)
prm = F.default_params(flags=flags)
cams = bench.camera_path(F, hgt, m, 512, 0, P, h, dist, stride=max(1, 512 // P))
arr = (F.Camera * P)(*cams)
dev = ctx.device_malloc(P * h * w * 4)
for _ in range(reps):
    ctx.render_batch_device(arr, prm, mp, h, w, dev)
ctx.sync()
ctx.set_profiling(True)
ctx.render_batch_device(arr, prm, mp, h, w, dev)
ctx.get_profile()
chunks, recs = ctx.get_counters()
nz = bench.n_z_of(F, prm, dist)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"workload": key, "poses": P, "flags": flags, "chunks_frac": chunks / (P * w * ((nz + 31) // 32)),
           "col_chunks_per_pose": chunks / P, "records_per_pose": recs / P},
          open("gpurun_out/prof_batch_%s_%d.json" % (key, P), "w"))
