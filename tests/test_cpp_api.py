"""The C++ face of the boundary: box2dpp_b200/include/Box2D/Gpu/b2WorldBatch.h keeps the reference's
b2World / b2BodyDef / b2FixtureDef / b2Shape spelling over the C ABI.  tests/cpp/hello_world_batch.cpp is the
reference's HelloWorld/HelloWorld.cpp:30-107 compiled against it (plus a multi-world pyramid batch with a
contact listener)."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_cases as G
from box2dpp_b200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "tests", "cpp", "hello_world_batch")


@pytest.fixture(scope="module")
def exe():
    lib = os.path.join(ROOT, "box2dpp_b200")
    if not os.path.exists(os.path.join(lib, "libb2gpu.so")):
        pytest.skip("libb2gpu.so not built")
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "box2dpp_b200", "include"), os.path.join(ROOT, "tests", "cpp", "hello_world_batch.cpp"),
           "-o", EXE, "-L" + lib, "-lb2gpu", "-Wl,-rpath," + lib]
    subprocess.check_call(cmd)
    return EXE


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_helloworld_compiles_and_refuses_cpu(exe):
    """The reference's HelloWorld source builds against the mirror API; without a GPU it fails loudly."""
    if _has_gpu():
        pytest.skip("CUDA device present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU path" in r.stderr


@pytest.mark.gpu
def test_helloworld_cpp_matches_golden(exe):
    """BASELINE config 1 through the C++ API: all 60 steps bit-identical to the reference (continuous physics off)."""
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [[int(x, 16) for x in line.split()] for line in r.stdout.strip().splitlines()]
    assert len(rows) == 60
    g = json.load(open(os.path.join(GOLD, "worlds.json")))["hello_world_noccd"]
    for step, row in enumerate(rows):
        data = np.array(row, dtype="<u4").tobytes()  # x, y, angle, vx, vy, w of the box
        assert "%016x" % G.fnv1a64(data) == g["state_hash"][step], f"step {step}"


@pytest.mark.gpu
def test_world_batch_cpp_with_listener(exe, oracles):
    """Three pyramids of different size in one b2WorldBatch, built through CreateBody / CreateFixture, with a
    b2ContactListener: event totals and the top box's state equal the oracle's."""
    n, rows, steps = 3, 4, 120
    r = subprocess.run([exe, "batch", str(n), str(rows), str(steps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.strip().splitlines()]
    assert len(lines) == n
    for w, f in enumerate(lines):
        sc = S.pyramid(rows + w, flags=G.NOCCD)
        ref = oracles.PortWorld(sc, record_events=True)
        counts = np.zeros(4, int)
        for _ in range(steps):
            ref.step(1.0 / 60.0, 8, 3)
            ei, _ = ref.events()
            counts += np.bincount(ei[:, 0], minlength=4)[:4]
        rb = ref.bodies()
        assert int(f[1]) == rb.shape[0]
        assert int(f[2]) == len(ref.contacts()[0])
        assert [int(x) for x in f[3:7]] == counts.tolist(), f"world {w}: listener event totals"
        top = rb[-1]
        got = [int(x, 16) for x in f[7:10]]
        want = [int(np.float32(v).view(np.uint32)) for v in (top[0], top[1], top[6])]
        assert got == want, f"world {w}: top box state"
        assert int(f[10]) == int(top[10] != 0)
