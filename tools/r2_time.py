"""Development timing: one workload of bench.py, device-resident frames, per-kernel split from the library's profiling
events.  usage: r2_time.py <workload> <poses> [flags] [reps]   (env FSB_MARCHC_VARIANT / FSB_SEGMENTS / FSB_MARCH_Z apply)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import futspace_b200 as F
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "1080p"]
P = int(sys.argv[2]) if len(sys.argv) > 2 else 256
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
ctx = F.Context(0)
col, hgt = F.terrain_fbm(m)
mp = ctx.upload_map(col, hgt)
prm = F.default_params(flags=flags)
cams = bench.camera_path(F, hgt, m, 512, 0, P, h, dist, stride=max(1, 512 // P))
arr = (F.Camera * P)(*cams)
dev = ctx.device_malloc(P * h * w * 4)
st = torch.cuda.ExternalStream(ctx.stream)
flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    if P == 1:
        ctx.render_device(cams[0], prm, mp, h, w, dev)
    else:
        ctx.render_batch_device(arr, prm, mp, h, w, dev)


for _ in range(3):
    step()
ctx.sync()
ms = bench.time_device_steps(torch, ctx, st, step, reps, flush)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(reps * 4):
    step()
e1.record(st)
e1.synchronize()
b2b = e0.elapsed_time(e1) / (reps * 4)
ctx.set_profiling(True)
step()
ctx.get_profile()
ctx.get_counters()
for _ in range(reps):
    step()
prof = ctx.get_profile()
chunks, recs = ctx.get_counters()
ctx.set_profiling(False)
nz = bench.n_z_of(F, prm, dist)
out = {"workload": sys.argv[1] if len(sys.argv) > 1 else "1080p", "poses": P, "flags": flags,
       "env": {k: os.environ.get(k) for k in ("FSB_MARCHC_VARIANT", "FSB_COLS_MIN_WARPS", "FSB_COLOUR_SLICE", "FSB_MARCH_Z", "FSB_GROUP_POSES", "FSB_FRAME_MAX_COLS", "FSB_EXPAND_TMA", "FSB_PDL", "FSB_SPLIT", "FSB_SPLIT_WARPS", "FSB_SPLIT_MAX_GROUPS", "FSB_LOCAL_CULL", "FSB_PAINT", "FSB_PAINT_SEG", "FSB_PAINT_VARIANT", "FSB_PAINT_PF", "FSB_MARCHC_PW", "FSB_FRAME_WARPS", "FSB_EXPAND_STAGE", "FSB_CARVEOUT", "FSB_PDL_BATCH") if os.environ.get(k)},
       "ms_per_step_flushed": sorted(ms)[len(ms) // 2], "ms_per_step_back_to_back": b2b,
       "frames_per_s": P / (sorted(ms)[len(ms) // 2] * 1e-3), "us_per_frame_b2b": 1e3 * b2b / P,
       "kernel_ms_per_step": {k: v[0] / reps for k, v in prof.items()},
       "paint_lane_utilisation": (recs / (32.0 * ctx.paint_trips)) if ctx.paint_trips else None,
       "chunks_frac": chunks / reps / (P * w * ((nz + 31) // 32)), "records_per_frame": recs / reps / P}
print(json.dumps(out), flush=True)
