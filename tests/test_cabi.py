"""CPU tests of the drop-in boundary: the library loads, exports every symbol include/*.h declares,
its host-side logic (z-series, parameter defaults, terrain generator) matches the oracle, and there
is no CPU fallback.  No compute entry point is called here (no GPU on the build box)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

try:
    import torch
    HAS_CUDA = torch.cuda.is_available()
except Exception:  # pragma: no cover
    HAS_CUDA = False


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            src = open(os.path.join(inc, fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b((?:fsb|futhark)_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(fsb):
    L = ctypes.CDLL(fsb.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 25
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing
    # and the Python binding covers the fsb_* part of the header
    assert set(n for n in decl if n.startswith("fsb_")) <= set(fsb.SYMBOLS)


def test_param_defaults_match_reference_constants(fsb, oracle):
    p, o = fsb.default_params(), oracle.default_params()
    assert (p.z0, p.delta, p.invz_param1, p.invz_param2) == (0.0, np.float32(0.001), 1.0, 0.0)
    assert (p.filter, p.sentinel, p.f2i_mode) == (o.filter, o.sentinel, o.f2i_mode) == (1, 0, 0)
    p, o = fsb.tests_variant_params(), oracle.tests_variant_params()
    assert (p.z0, p.delta, p.invz_param2, p.filter, p.sentinel) == (1.0, np.float32(0.005), 240.0, 0, 1)
    assert (o.z0, o.delta, o.invz_param2, o.filter, o.sentinel) == (1.0, np.float32(0.005), 240.0, 0, 1)


@pytest.mark.parametrize("delta,dist,z0", [(0.005, 800, 1), (0.001, 800, 0), (0.001, 1000, 0), (0.001, 2000, 0),
                                           (0.001, 4000, 0), (0.005, 4000, 1), (0.01, 123.4, 0.5), (0.001, 0.0004, 0)])
def test_get_zs_matches_oracle(fsb, oracle, delta, dist, z0):
    a, b = fsb.get_zs(delta, dist, z0), oracle.get_zs(delta, dist, z0)
    assert len(a) == len(b) and np.array_equal(a, b)


def test_get_zs_rejects_negative_distance(fsb):
    with pytest.raises(ValueError):
        fsb.get_zs(0.001, -4000.0, 0.0)


def test_terrain_is_deterministic_and_tileable(fsb):
    c1, h1 = fsb.terrain_fbm(512, seed=7)
    c2, h2 = fsb.terrain_fbm(512, seed=7)
    assert np.array_equal(c1, c2) and np.array_equal(h1, h2)
    assert h1.min() >= 0 and h1.max() <= 255 and (c1 >> 24 == 0xFF).all()
    # periodic: the wrap-around seam is as smooth as the interior
    seam = np.abs(h1[:, 0] - h1[:, -1]).max()
    interior = np.abs(np.diff(h1, axis=1)).max()
    assert seam <= interior + 1
    c3, _ = fsb.terrain_fbm(512, seed=8)
    assert not np.array_equal(c1, c3)
    with pytest.raises(fsb.FsbError):
        fsb.terrain_fbm(300)


@pytest.mark.skipif(HAS_CUDA, reason="checks the behaviour on a box without a GPU")
def test_no_cpu_fallback(fsb):
    with pytest.raises(fsb.FsbError) as e:
        fsb.Context(0)
    assert e.value.code == fsb.ERR_NO_DEVICE


@pytest.mark.parametrize("sh,sa", [(0.1, 0.1), (1.3, 0.4), (-0.7, 2.9), (0.0, 0.0)])
def test_sun_vector_matches_oracle(fsb, oracle, sh, sa):
    # vec3_rotate #y sun_ang (vec3_rotate #z sun_height [0,1,0]), fut/interactive.fut:56,196
    a, b = fsb.sun_vector(sh, sa), oracle.sun_vector(sh, sa)
    assert np.array_equal(np.float32(a), np.float32(b))
    assert abs(sum(x * x for x in a) - 1.0) < 1e-5


def test_headers_compile_as_c_and_cxx_and_link(fsb, tmp_path):
    """What a maintainer does first: include the two headers from C (c/interactive.c is C99-ish) and from C++, take the
    address of every declared entry point and link against the shared library (no call is made: no GPU needed)."""
    import subprocess
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    libdir = os.path.dirname(fsb.LIB_PATH)
    decl = declared_symbols()
    body = "\n".join("  p[%d] = (void (*)(void))%s;" % (i, n) for i, n in enumerate(decl))
    src = ('#include <stdio.h>\n#include "futspace_b200.h"\n#include "libfutspace.h"\n'
           "int main(void) {\n  void (*p[%d])(void);\n%s\n  printf(\"%%d\\n\", (int)(sizeof p / sizeof p[0]));\n  return p[0] == 0;\n}\n"
           % (len(decl), body))
    for name, cc, std in (("t.c", "/usr/bin/gcc", "-std=c99"), ("t.cpp", "/usr/bin/g++", "-std=c++11")):
        f = tmp_path / name
        f.write_text(src)
        exe = tmp_path / (name + ".out")
        subprocess.check_call([cc, std, "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", inc, str(f), "-o", str(exe),
                               "-L", libdir, "-lfutspace_b200", "-Wl,-rpath," + libdir])
        out = subprocess.check_output([str(exe)]).decode().strip()
        assert int(out) == len(decl)


def test_get_zs_random_agreement_with_oracle(fsb, oracle):
    """Library and oracle agree on the depth series for random constants -- values, length, and which inputs are rejected
    (negative z0 and tiny distances included)."""
    rng = np.random.default_rng(7)
    rejected = 0
    for _ in range(3000):
        delta = float(np.float32(rng.uniform(0.0005, 0.05)))
        z0 = float(np.float32(rng.uniform(-1, 3)))
        dist = float(np.float32(rng.choice([1e-5, 0.0011, 3.0, rng.uniform(0.01, 900)])))
        try:
            a = fsb.get_zs(delta, dist, z0)
        except ValueError:
            a = None
        try:
            b = oracle.get_zs(delta, dist, z0)
        except ValueError:
            b = None
        assert (a is None) == (b is None), (delta, dist, z0)
        if a is None:
            rejected += 1
        else:
            assert np.array_equal(a, b), (delta, dist, z0)
    assert 0 < rejected < 3000
