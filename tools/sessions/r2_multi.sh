#!/bin/bash
# Round-2 multi-GPU session (run under `gpurun --gpus N -- bash tools/sessions/r2_multi.sh N`): frame-parallel bench line and column split.
set -u
N=$1
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
if [ "${2:-all}" != "colsplit" ]; then
  timeout 900 $TR bench.py --gpus $N --steps 30 --warmup 3 ${3:-} > $O/bench_1080p_n$N.json 2> $O/bench_1080p_n$N.err; tail -c 400 $O/bench_1080p_n$N.err
  timeout 600 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/bench_reference_n$N.json 2> $O/bench_reference_n$N.err
fi
timeout 1200 $TR bench.py --gpus $N --workload 8k-colsplit --steps 10 > $O/bench_colsplit_n$N.json 2> $O/bench_colsplit_n$N.err; tail -c 400 $O/bench_colsplit_n$N.err
python - <<PY
import json
for f in ("bench_1080p_n$N", "bench_reference_n$N", "bench_colsplit_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f, "value", d.get("value"), "ms", d.get("ms_per_step"), "e2e", e.get("value"), "d2h GB/s", e.get("d2h_gb_per_s_aggregate"), "parity", d.get("parity_checked"), d.get("threads"))
        for k in ("fused_peer_store_ms_per_frame", "nccl_gather_ms_per_frame", "slab_render_ms_rank_max", "nvlink_ingest_gb_per_s_rank0"):
            if k in d: print("   ", k, d[k])
        if "configs" in d and "4k" in d["configs"]:
            c = d["configs"]["4k"]; print("    4k value", c["value"], "e2e", c["e2e"]["value"], c["parity_checked"])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
