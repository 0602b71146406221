"""sinf / cosf parity: the library's own b2Rot::Set arithmetic (box2dpp_b200/csrc/b2g_math.cuh SinCos,
shared by host layer and kernels) against the libm the reference links (glibc 2.39; frozen bit
patterns in tests/golden/sincos.npz, and the live libm when the reference oracle is present)."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_cases as G
from box2dpp_b200 import batch as B

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _call(fn, x, *pre):
    s, c = np.empty_like(x), np.empty_like(x)
    vp = C.c_void_p
    rc = fn(*pre, x.ctypes.data_as(vp), s.ctypes.data_as(vp), c.ctypes.data_as(vp), C.c_int64(x.size))
    assert rc == 0, B.lib().b2gpu_last_error()
    return s, c


def _same(a, b):
    a, b = a.view(np.uint32), b.view(np.uint32)
    nan = (np.isnan(a.view(np.float32)) & np.isnan(b.view(np.float32)))
    return bool(((a == b) | nan).all())


def test_host_sincos_matches_libm_golden():
    x = G.sincos_inputs()
    g = np.load(os.path.join(GOLD, "sincos.npz"))
    s, c = _call(B.lib().b2gpu_host_sincosf, x)
    assert _same(s, g["s"].view(np.float32)) and _same(c, g["c"].view(np.float32))


def test_wide_range_inline_sincos_matches_libm_golden_and_general_path():
    """MakeRotWideRange (|x| < 120, branch-free, what k_solve_position evaluates for turned bodies) == libm / SinCos."""
    x = G.sincos_inputs()
    g = np.load(os.path.join(GOLD, "sincos.npz"))
    s, c = _call(B.lib().b2gpu_host_sincosf_wide, x)
    assert _same(s, g["s"].view(np.float32)) and _same(c, g["c"].view(np.float32))
    rs = np.random.RandomState(5)
    y = np.concatenate([rs.uniform(-130, 130, 600000), rs.uniform(-1.0, 1.0, 100000), (np.arange(-160, 161) * (np.pi / 4)),
                        np.nextafter(np.float32(np.arange(-150, 151) * (np.pi / 4)), np.float32(0))]).astype(np.float32)
    s1, c1 = _call(B.lib().b2gpu_host_sincosf_wide, y)
    s2, c2 = _call(B.lib().b2gpu_host_sincosf, y)
    assert _same(s1, s2) and _same(c1, c2)


def test_host_sincos_matches_live_libm(oracles):
    if not oracles.available("ref"):
        pytest.skip("oracle/_ref/libb2ref.so not present on this box")
    rs = np.random.RandomState(77)
    x = np.concatenate([rs.uniform(-10, 10, 400000), rs.uniform(-1e4, 1e4, 100000)]).astype(np.float32)
    s, c = _call(B.lib().b2gpu_host_sincosf, x)
    rs_, rc_ = oracles.sincosf(x)
    assert _same(s, rs_) and _same(c, rc_)


@pytest.mark.gpu
def test_device_sincos_matches_libm_golden():
    x = G.sincos_inputs()
    g = np.load(os.path.join(GOLD, "sincos.npz"))
    s, c = _call(B.lib().b2gpu_sincosf, x, C.c_int32(0))
    assert _same(s, g["s"].view(np.float32)) and _same(c, g["c"].view(np.float32))
