"""bench.py's reference arm runs on the CPU alone: check the JSON-line contract without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "frames/s" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the same config keys as the GPU arm prints (bench.make_config), so the driver can compare the two lines
    for k in ("workload", "frame", "map", "distance", "n_z", "filter", "poses_per_gpu_per_step", "global_poses_per_step",
              "occlusion_bound", "parallelism", "l2"):
        assert k in d["config"], k


def test_reference_arm_ignores_omp_num_threads_and_never_loads_the_product_library():
    """torchrun exports OMP_NUM_THREADS=1 (round-1 SCALE run: the CPU arm fell to one thread at N >= 2); the arm passes
    its thread count explicitly.  It also generates the terrain from the oracle library, not from libfutspace_b200.so."""
    code = ("import sys, json, io, contextlib; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'cfg1', '--steps', '1', "
            "'--warmup', '0']; import bench; buf = io.StringIO();\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "d = json.loads(buf.getvalue()); maps = open('/proc/self/maps').read();\n"
            "print(json.dumps({'threads': d['threads'], 'cores': d['cpu_baseline']['cores'], "
            "'product_loaded': 'libfutspace_b200' in maps, 'oracle_loaded': 'libfs_oracle' in maps}))")
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr
    r = json.loads(p.stdout.strip().splitlines()[-1])
    assert r["threads"] == len(os.sched_getaffinity(0)) == r["cores"]
    assert r["oracle_loaded"] and not r["product_loaded"]


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0                                  # no CPU fallback for the product path
    assert not any(ln.startswith("{") for ln in p.stdout.splitlines())
