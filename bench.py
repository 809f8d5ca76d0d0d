#!/usr/bin/env python3
"""bench.py -- throughput of the futspace render hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k|cfg1] [--impl ours|reference]

One "step" = one camera-path batch of `poses` frames rendered through the C-ABI
(fsb_render_batch_device for `value`: frames stay in HBM; fsb_render_batch for `e2e`: pinned host
frames, H2D pose constants + D2H frames inside the timed region).  N > 1: one process per GPU
(torchrun), the global camera path is sharded frame-parallel, no data-path collective; time is the
max over ranks of the CUDA-event time on each rank's launch stream.

--impl reference times the CPU restatement of the reference (oracle/, OpenMP over columns, colour
evaluated for every sample as the reference does) on the host cores: the reference itself cannot be
built here (Futhark + un-vendored packages, DESIGN.md "Oracle").
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SKY = 0xFF9090E0
WORKLOADS = {
    # BASELINE.json configs[1] (+ configs[3]: the same frame as a 512-pose camera path)
    "1080p": dict(map=2048, w=1920, h=1080, dist=2000.0, poses=512, ref_poses=32,
                  name="1920x1080, 2048^2 synthetic fBm terrain, draw distance 2000, 512-pose camera path per GPU"),
    # BASELINE.json configs[2]
    "4k": dict(map=4096, w=3840, h=2160, dist=4000.0, poses=64, ref_poses=8,
               name="3840x2160, 4096^2 synthetic fBm terrain, draw distance 4000, 64-pose camera path per GPU"),
    # BASELINE.json configs[4]: one frame cut into column slabs, one slab per GPU, assembled on rank 0
    "8k-colsplit": dict(map=16384, w=7680, h=4320, dist=4000.0, poses=1, ref_poses=1, colsplit=True,
                        name="7680x4320 single frame, 16384^2 synthetic fBm terrain replicated per GPU, distance 4000, "
                             "column-split across the GPUs, slabs assembled in rank 0's frame"),
    # BASELINE.json configs[0] (synthetic stand-in for the converted map pair)
    "cfg1": dict(map=1024, w=1024, h=768, dist=1000.0, poses=512, ref_poses=64,
                 name="1024x768, 1024^2 synthetic fBm terrain, draw distance 1000, 512-pose camera path per GPU"),
}


def camera_path(F_or_O, height_map, m, n_total, first, count, h, dist, stride=1):
    """SURVEY.md 8d path; camera height clamped to terrain + 20 as terrain_collision would (fut/interactive.fut:67-87).
    Poses first, first + stride, ... (count of them) of an n_total-pose loop."""
    cams = []
    for i in range(first, first + count * stride, stride):
        th = 2.0 * math.pi * i / n_total
        x = m / 2 + (m / 4) * math.cos(th)
        y = m / 2 + (m / 4) * math.sin(th)
        ground = float(height_map[int(y) % m, int(x) % m])
        hgt = max(160.0 + 40.0 * math.sin(2 * th), ground + 20.0)
        cams.append(F_or_O.Camera(x, y, hgt, 2.2 + th, 0.3 * h, dist, 1.2, SKY))
    return cams


def n_z_of(F, prm, dist):
    return len(F.get_zs(prm.delta, dist, prm.z0))


class ClockSampler:
    """SM clock, power and throttle reasons sampled every 20 ms through NVML while the timed region runs
    (falls back to an `nvidia-smi -lms 100` subprocess when pynvml is not importable)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []
        self.stop_flag = False
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = sorted(x[0] for x in self.samples)
            reasons = sorted(k for k, b in bits.items() if any(x[2] & b for x in self.samples))
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.smax),
                    "power_w_max": max((x[1] for x in self.samples), default=None), "samples": len(sm),
                    "reasons": reasons, "source": "nvml, 20 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


def pin_to_gpu_numa_node(gpu_index):
    """Run this rank on the CPUs NVML reports as local to its GPU, so the pinned frame buffers of the e2e leg are
    first-touched on the GPU's own NUMA node (matters when several ranks stream frames to the host at once)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the march kernel from the committed ncu capture, if any (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


# ------------------------------------------------------------------------------------------------
def run_reference(args, wl):
    """CPU arm: the oracle (C restatement of the reference) on all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import futspace_b200 as F   # terrain generator + z-series only (host functions, no GPU)
    m, w, h, dist = wl["map"], wl["w"], wl["h"], wl["dist"]
    col, hgt = F.terrain_fbm(m)
    prm = O.default_params()
    n_ref = wl["ref_poses"]
    total = wl["poses"] * max(1, args.gpus)
    cores = os.cpu_count()
    # the sample: n_ref poses evenly spaced over the global path, a different offset each step
    def step(s):
        idx = [((s * 7 + j * (total // n_ref)) % total) for j in range(n_ref)]
        for i in idx:
            cam = camera_path(O, hgt, m, total, i, 1, h, dist)[0]
            O.render(cam, prm, col, hgt, h, w, eval_all_colors=True, nthreads=0)
    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s)
    dt = time.perf_counter() - t0
    fps = n_ref * args.steps / dt
    out = {
        "impl": "reference", "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "frame": [w, h], "map": m, "distance": dist, "filter": "bilinear",
                   "poses_per_step": n_ref},
        "mpixel_per_s": fps * w * h / 1e6,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "cpu_model": cpu_model(), "kind": "port",
                         "sample": "%d of the %d path poses per step, C restatement of the reference (oracle/), "
                                   "OpenMP over columns, colour filter evaluated for every sample" % (n_ref, total)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def time_device_steps(torch, ctx, st, fn, steps, flush):
    """-> list of per-step milliseconds (CUDA events on the launching stream); L2 flushed between steps."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        if flush is not None:
            with torch.cuda.stream(st):
                flush.add_(1)
        a.record(st)
        fn()
        b.record(st)
    ctx.sync()
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def run_ours(args, wl):
    import torch
    import futspace_b200 as F
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus):
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d (one process per GPU)" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    m, w, h, dst, P = args.map or wl["map"], wl["w"], wl["h"], wl["dist"], wl["poses"]
    if args.poses:
        P = args.poses
    col, hgt = F.terrain_fbm(m)
    ctx = F.Context(local)
    mp = ctx.upload_map(col, hgt)
    prm = F.default_params()
    nz = n_z_of(F, prm, dst)
    total = P * world
    cams = camera_path(F, hgt, m, total, rank, P, h, dst, stride=world)   # pose i -> rank i mod N (shard.pose_interleave)
    cam_arr = (F.Camera * P)(*cams)
    frame_bytes = w * h * 4
    st = torch.cuda.ExternalStream(ctx.stream)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    out_dev = ctx.device_malloc(P * frame_bytes)

    def step_dev():
        ctx.render_batch_device(cam_arr, prm, mp, h, w, out_dev)

    # ---- value: device-resident frames ----
    for _ in range(args.warmup):
        step_dev()
    barrier()
    ctx.sync()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ctx.launch_count
    ms = time_device_steps(torch, ctx, st, step_dev, args.steps, flush)
    launches = ctx.launch_count - n0
    clocks = sampler.stop()
    barrier()
    total_ms = max_over_ranks(sum(ms))
    value = total * args.steps / (total_ms / 1e3)

    # ---- the same steps with the occlusion bound off (every depth sample evaluated), for transparency ----
    prm_all = F.default_params(flags=F.FLAG_NO_CULL)

    def step_dev_all():
        ctx.render_batch_device(cam_arr, prm_all, mp, h, w, out_dev)

    for _ in range(3):
        step_dev_all()
    barrier()
    n_all = max(3, min(args.steps, 20))
    ms_all = time_device_steps(torch, ctx, st, step_dev_all, n_all, flush)
    value_all = total * n_all / (max_over_ranks(sum(ms_all)) / 1e3)

    # ---- per-kernel share + roofline of the dominant kernel (march) ----
    # The roofline is taken with the occlusion bound switched off (FSB_FLAG_NO_CULL): every depth sample is
    # fetched and evaluated, so the algorithmic bytes are really moved.  The default path (`value`) skips the
    # chunks the bound proves hidden; its kernel times and the fraction of chunks it evaluates are reported too.
    ctx.set_profiling(True)
    step_dev()
    ctx.get_profile()
    ctx.get_counters()
    for _ in range(2):
        step_dev()
    prof_cull = ctx.get_profile()
    chunks_eval, records = ctx.get_counters()
    prm_full = F.default_params(flags=F.FLAG_NO_CULL)

    def step_full():
        ctx.render_batch_device(cam_arr, prm_full, mp, h, w, out_dev)

    step_full()
    ctx.get_profile()
    ctx.get_counters()
    for _ in range(2):
        step_full()
    prof = ctx.get_profile()
    chunks_full, _ = ctx.get_counters()
    ctx.set_profiling(False)
    march_ms, march_n = prof["march"]
    expand_ms, expand_n = prof["expand"]
    setup_ms, setup_n = prof["setup"]
    poses_per_launch = 2.0 * P / march_n
    gather_bytes = 4.0 * 4 * w * nz                    # 4 taps x 4 B packed texel per sample
    frame_alg = 4.0 * w * h
    peak, peak_src = measured_peaks()
    achieved = gather_bytes * poses_per_launch / (march_ms / march_n * 1e-3) / 1e9
    tr = ncu_traffic(args.workload)
    tr = tr["dram_bytes_per_pose"] * poses_per_launch if tr else None
    roofline = {"bound": "hbm", "kernel": "fsb_march_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": tr, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": gather_bytes * poses_per_launch,
                "launch_ms": march_ms / march_n,
                "kernel_share_of_step": march_ms / (march_ms + expand_ms + setup_ms),
                "mode": "occlusion bound off (FSB_FLAG_NO_CULL): all W*n_z samples evaluated",
                "note": "frac can exceed 1: the algorithmic gather bytes (16 B per depth sample) are served from L1/L2 -- "
                        "neighbouring samples share 32-byte sectors of the 2-byte height texture -- so HBM is the nominal "
                        "bound the contract asks for, not the limiter; ncu (profiles/r1_end_march_*) shows the kernel "
                        "issue-bound: 76 % issue-active, XU pipe 60 %, DRAM 3.5 %",
                "default_path": {"march_launch_ms": prof_cull["march"][0] / prof_cull["march"][1],
                                 "chunks_evaluated_frac": chunks_eval / max(1, chunks_full),
                                 "records_per_frame": records / (2.0 * P),
                                 "kernel_ms_per_step": {k: v[0] / 2 for k, v in prof_cull.items()}}}
    step_alg = (gather_bytes + frame_alg) * P
    full_ms = (march_ms + expand_ms + setup_ms) / 2
    roofline_step = {"achieved": step_alg / (full_ms * 1e-3) / 1e9, "unit": "GB/s per GPU (occlusion bound off)",
                     "frac": step_alg / (full_ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_step_per_gpu": step_alg,
                     "kernel_ms_per_step": {"setup": setup_ms / 2, "march": march_ms / 2, "expand": expand_ms / 2}}

    # ---- e2e: host frames through fsb_render_batch (pinned), copies inside the timed region ----
    host = ctx.host_malloc(P * frame_bytes)

    def step_e2e():
        ctx.render_batch(cam_arr, prm, mp, h, w, out=host)

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        step_e2e()
    ctx.sync()
    e_dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total * e_steps / e_dt
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": 64 * P, "d2h_bytes_per_step": P * frame_bytes,
           "steps": e_steps, "ms_per_step": 1e3 * e_dt / e_steps,
           "api": "fsb_render_batch (pinned host frames; pose constants H2D + frames D2H inside the timed region)"}

    # ---- secondary runs: other renderer variants on the same path, single-frame latency ----
    def timed(prm_x, n_steps=5):
        def f():
            ctx.render_batch_device(cam_arr, prm_x, mp, h, w, out_dev)
        f()
        msx = time_device_steps(torch, ctx, st, f, n_steps, flush)
        return total * n_steps / (max_over_ranks(sum(msx)) / 1e3)

    secondary = {
        "tests_variant_frames_per_s": timed(F.tests_variant_params()),   # z0 = 1, d = 0.005, nearest, sky sentinel, 240
        "smoothing_on_frames_per_s": timed(F.default_params(flags=F.FLAG_SMOOTHING)),
        "nearest_frames_per_s": timed(F.default_params(filter=0)),
    }
    single = F.Camera(m / 2 + 0.37, m / 2 + 0.73, max(160.0, float(hgt[m // 2, m // 2]) + 20.0), 2.2, 0.3 * h, dst, 1.2, SKY)

    def one_frame():
        ctx.render_device(single, prm, mp, h, w, out_dev)

    one_frame()
    ms1 = time_device_steps(torch, ctx, st, one_frame, 20, flush)
    secondary["single_frame_cold_l2_us"] = 1e3 * sorted(ms1)[len(ms1) // 2]   # median; 3 launches, L2 flushed before each
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50):
        one_frame()
    e1.record(st)
    e1.synchronize()
    secondary["single_frame_back_to_back_us"] = 1e3 * e0.elapsed_time(e1) / 50   # map L2-resident, launches queued

    if args.workload == "cfg1":
        # BASELINE config 1 names the reference's own converted map pair: C1W / D1 (committed copy, tests/golden/)
        import numpy as np
        gold = os.path.join(ROOT, "tests", "golden", "c1w_d1.npz")
        if os.path.exists(gold):
            z = np.load(gold)
            rgb = (z["r"].astype(np.uint32) << 16) | (z["g"].astype(np.uint32) << 8) | z["b"].astype(np.uint32) | np.uint32(0xFF000000)
            h2 = z["height"].astype(np.int32)
            mp2 = ctx.upload_map(rgb, h2)
            cams2 = camera_path(F, h2, 1024, total, rank, P, h, dst, stride=world)
            arr2 = (F.Camera * P)(*cams2)

            def f2():
                ctx.render_batch_device(arr2, prm, mp2, h, w, out_dev)
            f2()
            ms2 = time_device_steps(torch, ctx, st, f2, 5, flush)
            secondary["c1w_d1_map_frames_per_s"] = total * 5 / (max_over_ranks(sum(ms2)) / 1e3)
            mp2.free()
    extra = {"secondary": secondary}
    cpu = None
    if rank == 0:
        extra["l2_stream_gbs"] = ctx.l2_stream_gbs()
        extra["l2_gather_gsectors"] = ctx.l2_gather_gsectors()
        taps_per_s = 4.0 * w * nz * poses_per_launch / (march_ms / march_n * 1e-3)
        extra["l2_gather_roofline_frac"] = taps_per_s / (extra["l2_gather_gsectors"] * 1e9)
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(F, wl, col, hgt, total)
    ctx.host_free(host)
    ctx.device_free(out_dev)
    mp.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    out = {
        "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "frame": [w, h], "map": m, "distance": dst, "n_z": nz,
                   "filter": "bilinear", "poses_per_gpu_per_step": P, "global_poses_per_step": total,
                   "occlusion_bound": "on (default path; frames bit-identical to the full evaluation, see roofline.default_path)",
                   "parallelism": "frame-parallel x%d (pose i -> GPU i mod N), maps replicated, no collective" % world,
                   "l2": "256 MiB scratch write between timed steps (flush); each step also streams %.1f GB of "
                         "frames through the 126 MB L2; the %d MiB packed map is L2-resident by design"
                         % (P * frame_bytes / 1e9, m * m * 4 >> 20)},
        "mpixel_per_s": value * w * h / 1e6,
        "value_full_evaluation": value_all,   # occlusion bound off: all W*n_z samples fetched and evaluated
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_step": roofline_step,
        "cpu_baseline": cpu,
    }
    out.update(extra)
    print(json.dumps(out), flush=True)


def run_colsplit(args, wl):
    """Column-split mode: rank r renders columns [b[r], b[r+1]) of ONE frame.  Two ways to assemble it on
    rank 0 are timed: (a) fused -- every rank's expand kernel stores straight into rank 0's frame through a CUDA-IPC
    peer mapping (NVLink), no separate collective; (b) each rank renders a private slab and NCCL gathers them."""
    import numpy as np
    import torch
    import futspace_b200 as F
    from futspace_b200.shard import column_bounds, gather_columns
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    m = args.map or wl["map"]
    w, h, dst = wl["w"], wl["h"], wl["dist"]
    col, hgt = F.terrain_fbm(m)
    ctx = F.Context(local)
    mp = ctx.upload_map(col, hgt)
    prm = F.default_params()
    nz = n_z_of(F, prm, dst)
    cam = F.Camera(m / 2 + 0.37, m / 2 + 0.73, max(160.0, float(hgt[m // 2, m // 2]) + 20.0), 2.2, 0.3 * h, dst, 1.2, SKY)
    del col
    b = column_bounds(w, world)
    c0, c1 = b[rank], b[rank + 1]
    st = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # (a) fused: peer stores into rank 0's frame
    frame = ctx.device_malloc(h * w * 4) if rank == 0 else None
    handle = [ctx.ipc_export(frame) if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(handle, src=0)
    base = frame if rank == 0 else ctx.ipc_import(handle[0])

    def step_fused():
        ctx.render_columns_device(cam, prm, mp, h, w, c0, c1, base + 4 * c0, w)

    def timed(fn, steps):
        out = []
        for _ in range(steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            fn()
            e1.record(st)
            e1.synchronize()
            out.append(e0.elapsed_time(e1))
        return reduce_max(out)   # a frame is complete when the slowest slab has landed

    for _ in range(args.warmup):
        step_fused()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = ctx.launch_count
    fused = timed(step_fused, args.steps)
    launches = ctx.launch_count - n0
    clocks = sampler.stop()
    barrier()
    ok = None
    if rank == 0:   # the assembled frame equals a single-GPU render of the whole frame
        ref = ctx.device_malloc(h * w * 4)
        ctx.render_device(cam, prm, mp, h, w, ref)
        ok = bool(np.array_equal(ctx.download(frame, (h, w)), ctx.download(ref, (h, w))))
        ctx.device_free(ref)

    # (b) private slabs + NCCL gather to rank 0
    gather_ms = None
    if dist is not None:
        wmax = max(b[i + 1] - b[i] for i in range(world))
        slab = torch.zeros((h, wmax), dtype=torch.int32, device="cuda")

        def step_gather():
            ctx.render_columns_device(cam, prm, mp, h, w, c0, c1, slab.data_ptr(), wmax)
            ctx.sync()
            return gather_columns(dist, slab, b, h, w, rank, world)

        for _ in range(2):
            step_gather()
        tt = []
        for _ in range(max(1, min(args.steps, 10))):
            barrier()
            t0 = time.perf_counter()
            step_gather()
            torch.cuda.synchronize()
            tt.append(1e3 * (time.perf_counter() - t0))
        gather_ms = sum(reduce_max(tt)) / len(tt)
    if rank != 0 and dist is not None:
        ctx.ipc_close(base)
    barrier()
    if rank == 0:
        ms = sum(fused) / len(fused)
        peak, peak_src = measured_peaks()
        alg = 4.0 * 4 * w * nz + 4.0 * w * h
        print(json.dumps({
            "metric": "frames/s", "value": 1e3 / ms, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "frame": [w, h], "map": m, "distance": dst, "n_z": nz,
                       "column_bounds": b, "parallelism": "column-split x%d, maps replicated; slabs stored into rank 0's "
                       "frame through CUDA-IPC peer mappings (NVLink), no separate collective" % world,
                       "l2": "single frame per step; the touched map footprint (~77 MB packed at distance 4000) and the "
                             "133 MB frame exceed what stays resident between steps"},
            "mpixel_per_s": w * h / ms / 1e3, "clocks": clocks, "gpu_launches": launches,
            "frame_matches_single_gpu": ok,
            "fused_peer_store_ms_per_frame": ms, "nccl_gather_ms_per_frame": gather_ms,
            "roofline_step": {"achieved": alg / (ms * 1e-3) / 1e9, "unit": "GB/s (algorithmic, whole frame)", "frac": alg / (ms * 1e-3) / 1e9 / peak / world,
                              "peak_source": peak_src},
            "e2e": None, "cpu_baseline": None,
        }), flush=True)
    mp.free()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(F, wl, col, hgt, total):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    m, w, h, dst = wl["map"], wl["w"], wl["h"], wl["dist"]
    prm = O.default_params()
    cam0 = camera_path(O, hgt, m, total, 0, 1, h, dst)[0]
    O.render(cam0, prm, col, hgt, h, w, eval_all_colors=True, nthreads=0)   # warm-up
    # bounded sample: poses spread over the path (stride 37 is coprime to the path length) until ~12 s of CPU work
    t0 = time.perf_counter()
    n = 0
    while n < total and time.perf_counter() - t0 < 12.0:
        i = (n * 37) % total
        O.render(camera_path(O, hgt, m, total, i, 1, h, dst)[0], prm, col, hgt, h, w, eval_all_colors=True, nthreads=0)
        n += 1
    dt = time.perf_counter() - t0
    # the `futhark c` (sequential backend) analogue: the same code on one thread, two poses
    t1 = time.perf_counter()
    for i in (0, total // 2):
        O.render(camera_path(O, hgt, m, total, i, 1, h, dst)[0], prm, col, hgt, h, w, eval_all_colors=True, nthreads=1)
    one = 2.0 / (time.perf_counter() - t1)
    return {"value": n / dt, "unit": "frames/s", "cores": os.cpu_count(), "cpu_model": cpu_model(), "kind": "port",
            "one_thread_value": one,
            "sample": "%d poses spread over the %d-pose path (%.1f s), C restatement of the reference (oracle/), "
                      "OpenMP over columns, colour filter evaluated for every sample as the reference does; "
                      "one_thread_value = 2 poses on a single thread" % (n, total, dt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=0, help="override poses per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--map", type=int, default=0, help="override the map size (multiple of 256)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    elif wl.get("colsplit"):
        run_colsplit(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
