import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


SKY = 0xFF9090E0


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def fsb():
    """The product binding.  Builds the library if the in-tree .so is missing (nvcc cross-compiles on CPU)."""
    import futspace_b200 as F
    if not os.path.exists(F.LIB_PATH):
        F.build()
    F.lib()
    return F


@pytest.fixture(scope="session")
def c1w_d1():
    """The reference's C1W/D1 map pair (committed copy, see tests/golden/make_golden.py)."""
    z = np.load(os.path.join(HERE, "golden", "c1w_d1.npz"))
    rgb = (z["r"].astype(np.uint32) << 16) | (z["g"].astype(np.uint32) << 8) | z["b"].astype(np.uint32)
    return rgb, z["height"].astype(np.int32)


@pytest.fixture(scope="session")
def real_maps(c1w_d1):
    """Four of the reference's 29 map pairs (C1W/D1 plus tests/golden/make_golden_maps.py): name -> (rgb u32, height i32)."""
    maps = {"C1W/D1": c1w_d1}
    for n in (7, 13, 29):
        z = np.load(os.path.join(HERE, "golden", "c%dw_d%d_pal.npz" % (n, n)))
        maps["C%dW/D%d" % (n, n)] = (z["pal"][z["idx"]].astype(np.uint32), z["height"].astype(np.int32))
    return maps


@pytest.fixture(scope="session")
def golden_frames():
    return np.load(os.path.join(HERE, "golden", "golden_frames.npz"))


@pytest.fixture(scope="session")
def fbm1024(fsb):
    return fsb.terrain_fbm(1024)


@pytest.fixture(scope="session")
def gpu_ctx(fsb):
    ctx = fsb.Context(0)   # raises on a box without an sm_100 GPU: there is no fallback
    yield ctx
    ctx.close()
