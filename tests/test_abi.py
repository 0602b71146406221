"""The drop-in boundary: libb2gpu.so loads, exports every entry point include/b2gpu.h declares, and
fails loudly (never falls back to a CPU path) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from box2dpp_b200 import batch as B
from box2dpp_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2gpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = B.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/b2gpu.h but not exported by libb2gpu.so"
    for n in B.EXPORTED_SYMBOLS:
        assert n in names, f"binding lists {n} which the header does not declare"


def test_scene_header_matches_ctypes_mirror():
    """sizeof of the ctypes mirrors == the C structs (checked through b2gpu_scene_capacities reading them)."""
    sc = S.pyramid(3)
    d = B.BatchDesc()
    rc = B.lib().b2gpu_scene_capacities(C.byref(sc.c_scene()), 1, C.byref(d))
    assert rc == 0
    assert d.maxBodies >= 7 and d.maxFixtures >= 7 and d.maxProxies >= 7 and d.maxContacts > 0 and d.maxShapeWords > 0


def test_joint_defs_reach_the_host_layer():
    """b2g_joint_def / the joint fields of b2g_scene as ctypes sees them == as the C side reads them."""
    sc = S.jointed(1)
    d = B.BatchDesc()
    assert B.lib().b2gpu_scene_capacities(C.byref(sc.c_scene()), 1, C.byref(d)) == 0
    assert d.maxJoints == len(sc.joints) == 22 and d.maxBodies == len(sc.bodies)
    assert d.reserved & 4  # circles + boxes + an edge: the type-binned narrow phase is selected
    d2 = B.BatchDesc()
    assert B.lib().b2gpu_scene_capacities(C.byref(S.pyramid(3).c_scene()), 1, C.byref(d2)) == 0
    assert d2.maxJoints == 0 and (d2.reserved & 4) == 0


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_no_cpu_fallback():
    if _has_gpu():
        pytest.skip("CUDA device present")
    with pytest.raises(B.B2GpuError) as e:
        B.WorldBatch([S.hello_world()])
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)
    x = np.zeros(4, np.float32)
    vp = C.c_void_p
    rc = B.lib().b2gpu_sincosf(0, x.ctypes.data_as(vp), x.ctypes.data_as(vp), x.ctypes.data_as(vp), C.c_int64(4))
    assert rc == -1  # B2GPU_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under box2dpp_b200/ or include/ may reference it."""
    for base in ("box2dpp_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".h", ".cuh", ".cu", ".cpp", ".hpp")):
                    text = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle/|libb2oracle|libb2ref", text, re.M), \
                        f"{dp}/{f} references the oracle"
