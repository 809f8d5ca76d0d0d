"""The reference's test pipeline (tests/Makefile:1-5): `cat C1W.in D1.in | ./futspace -b > img_map.data`."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SKY
from futspace_b200 import futdata

CLI = os.path.join(ROOT, "futspace_b200", "futspace")


def test_futhark_binary_format_roundtrip():
    a = np.arange(12, dtype=np.int32).reshape(3, 4) - 5
    buf = futdata.dumps(a)
    assert buf[:7] == b"b\x02\x02 i32" and len(buf) == 7 + 16 + 48     # tools/png2data.py:50-57
    b, end = futdata.loads(buf)
    assert end == len(buf) and np.array_equal(a, b) and b.dtype == np.int32
    two = buf + b"\n" + futdata.dumps(a.astype(np.uint32))
    x, off = futdata.loads(two)
    y, _ = futdata.loads(two, off)
    assert y.dtype == np.uint32 and np.array_equal(y, a.astype(np.uint32))


def test_cli_fails_loudly_without_gpu(fsb):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    assert os.path.exists(CLI)
    p = subprocess.run([CLI, "-b"], input=b"", capture_output=True)
    assert p.returncode == 1 and b"no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_cli_reproduces_tests_variant_frame(fsb, c1w_d1, golden_frames, tmp_path):
    rgb, hgt = c1w_d1
    stdin = futdata.dumps(rgb.astype(np.int32)) + futdata.dumps(hgt.astype(np.int32))   # C1W.in, D1.in
    p = subprocess.run([CLI, "-b", "-D", "-t", str(tmp_path / "t.txt")], input=stdin, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()
    frame, _ = futdata.loads(p.stdout)
    assert frame.shape == (400, 800) and frame.dtype == np.int32
    assert np.array_equal(frame.view(np.uint32), golden_frames["tests_variant_400x800"])
    assert int(open(tmp_path / "t.txt").read()) > 0 and b"device" in p.stderr
    # text output (no -b) carries the same numbers
    p = subprocess.run([CLI], input=stdin, capture_output=True)
    assert p.returncode == 0
    txt = p.stdout.decode()
    assert txt.startswith("[[") and txt.count("i32") == 400 * 800
    first = int(txt[2:txt.index("i32")])
    assert first == int(frame[0, 0])
