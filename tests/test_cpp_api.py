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


@pytest.fixture(scope="module")
def edits_exe():
    lib = os.path.join(ROOT, "box2dpp_b200")
    if not os.path.exists(os.path.join(lib, "libb2gpu.so")):
        pytest.skip("libb2gpu.so not built")
    out = os.path.join(ROOT, "tests", "cpp", "scene_open_edits")
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "box2dpp_b200", "include"), os.path.join(ROOT, "tests", "cpp", "scene_open_edits.cpp"),
           "-o", out, "-L" + lib, "-lb2gpu", "-Wl,-rpath," + lib]
    subprocess.check_call(cmd)
    return out


def _build(name):
    lib = os.path.join(ROOT, "box2dpp_b200")
    if not os.path.exists(os.path.join(lib, "libb2gpu.so")):
        pytest.skip("libb2gpu.so not built")
    out = os.path.join(ROOT, "tests", "cpp", name)
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "box2dpp_b200", "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-o", out, "-L" + lib, "-lb2gpu", "-Wl,-rpath," + lib]
    subprocess.check_call(cmd)
    return out


def test_world_edits_program_compiles():
    """The running-world edit calls of the mirror (CreateJoint, CreateFixture on an existing body, DestroyFixture, SetType ...)
    compile warning-free against the header."""
    _build("world_edits")


@pytest.mark.gpu
def test_running_world_creation_equals_scene_time_creation():
    """A fixture and a joint given to a running world == the same calls made before the first Step, bit for bit after 90
    steps; DestroyJoint / DestroyFixture / SetMassData / SetActive / SetType / SetLinearVelocity behave (tests/cpp/world_edits.cpp)."""
    r = subprocess.run([_build("world_edits")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    rows = r.stdout.strip().splitlines()
    assert len(rows) == 3 and rows[0] == rows[1], rows
    assert rows[2] == "edits ok", rows


def test_edits_and_getters_before_the_first_step_do_not_freeze_the_scene(edits_exe):
    """b2Body setters / getters right after CreateBody and joint defs initialised from GetWorldCenter() (the reference's
    usual construction patterns) are served on the host: no device batch, the scene stays open.  Needs no GPU."""
    r = subprocess.run([edits_exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "no device batch created" in r.stdout


@pytest.mark.gpu
def test_setters_before_the_first_step_equal_the_definition(edits_exe):
    """SetLinearVelocity / SetAngularVelocity after CreateBody == the same values in the b2BodyDef, bit for bit after 60 steps;
    CreateJoint on the running world throws instead of returning nullptr."""
    r = subprocess.run([edits_exe, "step"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = r.stdout.strip().splitlines()
    assert len(rows) == 2 and rows[0] == rows[1], rows


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


@pytest.mark.gpu
def test_body_edits_through_cpp_api(exe, oracles):
    """b2Body::ApplyLinearImpulse / ApplyForce / SetTransform / SetAwake ... called between steps through the C++
    mirror (queued as b2gpu_body_command): final state of every body equals the oracle's, bit for bit."""
    from box2dpp_b200 import batch as B
    n, steps = 3, 150
    r = subprocess.run([exe, "edits", str(n), str(steps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {}
    for line in r.stdout.strip().splitlines():
        f = line.split()
        got[(int(f[0]), int(f[1]))] = [int(x, 16) for x in f[2:8]] + [int(f[8])]
    Ref = oracles.RefWorld if oracles.available("ref") else oracles.PortWorld
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    for w in range(n):
        ref = Ref(S.pyramid(4, flags=G.NOCCD))
        nb = ref.bodies().shape[0]
        for s in range(steps):
            t = 1 + (s + w) % (nb - 1)
            k = (s + 2 * w) % 7
            if k == 0:
                c = ref.bodies()[t]
                ref.command(t, B.CMD_APPLY_LINEAR_IMPULSE, 1, (3.0 + w, 12.0), (float(c[4]), float(c[5])))
            elif k == 1:
                ref.command(t, B.CMD_APPLY_FORCE, s & 1, (-150.0, 400.0), (f32(0.25 * s), 2.0))
            elif k == 2:
                ref.command(t, B.CMD_SET_ANGULAR_VELOCITY, 1, (f32(0.5 * w - 1.0), 0.0))
            elif k == 3:
                if s % 21 == 3:
                    ref.command(t, B.CMD_SET_TRANSFORM, 1, (-2.0 + w, 6.0), (f32(0.3), 0.0))
            elif k == 4:
                ref.command(t, B.CMD_APPLY_TORQUE, 1, (25.0, 0.0))
                ref.command(t, B.CMD_APPLY_FORCE_TO_CENTER, 1, (40.0, 0.0))
            elif k == 5:
                if s % 35 == 5:
                    ref.command(t, B.CMD_SET_AWAKE, 0)
                else:
                    ref.command(t, B.CMD_APPLY_ANGULAR_IMPULSE, 1, (-1.5, 0.0))
            else:
                ref.command(t, B.CMD_SET_LINEAR_VELOCITY, 1, (0.0, 4.0 if s % 14 == 6 else 0.0))
            ref.step(1.0 / 60.0, 8, 3)
        rb = ref.bodies()
        for i in range(nb):
            want = [int(np.float32(v).view(np.uint32)) for v in (rb[i, 0], rb[i, 1], rb[i, 6], rb[i, 7], rb[i, 8], rb[i, 9])]
            assert got[(w, i)][:6] == want, f"world {w} body {i}"
            assert got[(w, i)][6] == int(rb[i, 10] != 0)


@pytest.mark.gpu
def test_joints_through_cpp_api(exe, oracles):
    """b2PrismaticJointDef / b2RevoluteJointDef / b2DistanceJointDef with Initialize() and b2World::CreateJoint through
    the C++ mirror (reference defaults: continuous physics on): state after 200 steps equals the reference's."""
    if not oracles.available("ref"):
        pytest.skip("joints are checked against the live reference")
    r = subprocess.run([exe, "joints", "200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [[int(x, 16) for x in line.split()] for line in r.stdout.strip().splitlines()]
    s = S.Scene()
    g = s.create_body(S.STATIC_BODY)
    s.create_edge(g, (-40.0, 0.0), (40.0, 0.0))
    cart = s.create_body(S.DYNAMIC_BODY, (0.0, 3.0))
    s.create_box(cart, 0.6, 0.25, density=2.0)
    slide = s.create_prismatic_joint(g, cart, (0.0, 3.0), (1.0, 0.0), lower=-3.0, upper=3.0, enable_limit=True,
                                     max_motor_force=60.0, motor_speed=2.5, enable_motor=True)
    pole = s.create_body(S.DYNAMIC_BODY, (0.0, 4.0), angle=0.2)
    s.create_box(pole, 0.08, 1.0, density=1.0)
    s.create_revolute_joint(cart, pole, (0.0, 3.0))
    bob = s.create_body(S.DYNAMIC_BODY, (8.0, 6.0))
    s.create_circle(bob, 0.35, density=2.0)
    s.create_distance_joint(g, bob, (6.0, 8.0), (8.0, 6.0), frequency_hz=3.0, damping_ratio=0.2)
    wheel = s.create_body(S.DYNAMIC_BODY, (-6.0, 5.0))
    s.create_circle(wheel, 0.5, density=3.0)
    wheel_joint = s.create_revolute_joint(g, wheel, (-6.0, 5.0))
    s.create_gear_joint(wheel_joint, slide, ratio=-2.0)
    crate = s.create_body(S.DYNAMIC_BODY, (12.0, 4.0))
    s.create_box(crate, 0.4, 0.4, density=1.0)
    s.create_mouse_joint(g, crate, (12.2, 4.3), max_force=500.0)
    ref = oracles.RefWorld(s)
    for _ in range(200):
        ref.step(1.0 / 60.0, 8, 3)
    rb = ref.bodies()
    for i, b in enumerate((cart, pole, bob, wheel, crate)):
        want = [int(np.float32(v).view(np.uint32)) for v in (rb[b, 0], rb[b, 1], rb[b, 6])]
        assert got[i] == want, f"body {b}: {got[i]} vs {want}"


@pytest.mark.gpu
def test_spawn_and_despawn_through_cpp_api(exe, oracles):
    """b2World::CreateBody / b2Body::CreateFixture / b2World::DestroyBody on running worlds through the C++ mirror
    (reference defaults): every surviving body's state equals the reference's after 160 steps."""
    if not oracles.available("ref"):
        pytest.skip("checked against the live reference")
    from box2dpp_b200 import batch as B
    steps = 160
    r = subprocess.run([exe, "spawn", str(steps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {}
    for line in r.stdout.strip().splitlines():
        f = line.split()
        got[(int(f[0]), int(f[1]))] = None if f[2] == "dead" else [int(x, 16) for x in f[2:5]]
    f32 = lambda x: float(np.float32(x))  # noqa: E731
    for w in range(2):
        s = S.Scene()
        g = s.create_body(S.STATIC_BODY)
        s.create_box(g, 20.0, 1.0, center=(0.0, -1.0), angle=0.0)
        b = s.create_body(S.DYNAMIC_BODY, (0.0, 2.0 + w))
        s.create_box(b, 0.5, 0.5, density=1.0)
        ref = oracles.RefWorld(s)
        count = 2
        for st in range(steps):
            if st % 20 == 10 and count < 8:
                k = st // 20
                tmp = S.Scene()
                px = f32(f32(0.3) * np.float32(k) - f32(0.4) * np.float32(w))
                nb = tmp.create_body(S.DYNAMIC_BODY, (px, 5.0), angle=f32(f32(0.1) * np.float32(k)))
                if k % 2 == 0:
                    tmp.create_box(nb, 0.4, 0.3, density=2.0, friction=0.5)
                else:
                    tmp.create_circle(nb, 0.35, density=1.5)
                idx = ref.create_body(tmp.bodies[0], list(tmp.fixtures), list(tmp.vertices))
                ref.command(idx, B.CMD_APPLY_LINEAR_IMPULSE, 1, (0.5, 0.0), (px, f32(5.1)))
                count += 1
            if st == 55 + 20 * w:
                ref.command(1, B.CMD_DESTROY_BODY)
            ref.step(1.0 / 60.0, 8, 3)
        rb = ref.bodies()
        assert rb.shape[0] == count
        for i in range(count):
            if rb[i, 15] == 1.0:
                assert got[(w, i)] is None
                continue
            want = [int(np.float32(v).view(np.uint32)) for v in (rb[i, 0], rb[i, 1], rb[i, 6])]
            assert got[(w, i)] == want, f"world {w} body {i}: {got[(w, i)]} vs {want}"
