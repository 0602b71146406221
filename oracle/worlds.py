"""TEST INFRASTRUCTURE — ctypes wrappers over the two CPU oracles:

  * ``RefWorld``  -> oracle/_ref/libb2ref.so: the reference's own vanilla Box2D 2.3.2, compiled from
    /root/reference by oracle/Makefile behind oracle/ref_harness.cpp ("reference");
  * ``PortWorld`` -> oracle/libb2oracle.so: the plain-C restatement in oracle/b2o_*.c ("port").

Both expose the same calls and dump formats.  Only tests/, bench.py's cpu_baseline / --impl reference
leg and __graft_entry__.smoke() may import this module; the product never does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from box2dpp_b200.scene import CScene, FixtureDef, Scene, scene_array

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libb2ref.so")
PORT_SO = os.path.join(_HERE, "libb2oracle.so")

_libs = {}


def available(kind="ref") -> bool:
    return os.path.exists(REF_SO if kind == "ref" else PORT_SO)


class _Prefixed:
    """Maps L.ref_xxx to the same-named b2o_xxx symbol of the port library."""

    def __init__(self, cdll, prefix):
        self._cdll, self._prefix = cdll, prefix

    def __getattr__(self, name):
        if name.startswith("ref_"):
            name = self._prefix + name[4:]
        return getattr(self._cdll, name)


def lib(kind="ref"):
    if kind not in _libs:
        if kind == "ref":
            L = C.CDLL(REF_SO)
        else:
            L = _Prefixed(C.CDLL(PORT_SO), "b2o_")
        L.ref_world_create.restype = C.c_void_p
        L.ref_world_create.argtypes = [C.POINTER(CScene)]
        L.ref_world_destroy.argtypes = [C.c_void_p]
        L.ref_world_record_events.argtypes = [C.c_void_p, C.c_int]
        L.ref_world_step.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        for f in ("ref_world_body_count", "ref_world_contact_count", "ref_world_proxy_count"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.ref_world_get_bodies.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_world_get_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_world_get_contacts.restype = C.c_int
        L.ref_world_get_proxies.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_world_get_proxies.restype = C.c_int
        L.ref_world_get_events.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_world_get_events.restype = C.c_int
        if kind == "ref":
            L.ref_world_create_body.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_world_create_body.restype = C.c_int
            L.ref_world_set_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
            L.ref_world_set_filter.restype = None
            L.ref_world_set_density.argtypes = [C.c_void_p, C.c_int, C.c_float]
            L.ref_world_set_density.restype = None
            L.ref_world_fixture_mass_data.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
            L.ref_world_fixture_mass_data.restype = C.c_int
            L.ref_world_create_joint.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_world_create_joint.restype = C.c_int
            L.ref_world_create_fixture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
            L.ref_world_create_fixture.restype = C.c_int
        L.ref_world_body_command.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.ref_world_body_command.restype = None
        L.ref_polygon_build.argtypes = [C.POINTER(FixtureDef), C.c_void_p, C.c_void_p]
        L.ref_polygon_build.restype = C.c_int
        for f in ("ref_collide_polygons", "ref_collide_circles", "ref_collide_polygon_circle",
                  "ref_collide_edge_circle", "ref_collide_edge_polygon"):
            getattr(L, f).argtypes = [C.c_void_p] * 6
        if kind == "ref":
            L.ref_world_query_aabb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
            L.ref_world_query_aabb.restype = C.c_int
            L.ref_world_raycast_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.ref_world_raycast_closest.restype = C.c_int
            L.ref_sincosf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
            L.ref_bench_run.restype = C.c_double
            L.ref_bench_run.argtypes = [C.POINTER(CScene), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_int, C.c_int, C.POINTER(C.c_uint64)]
            L.ref_bench_run_repeats.restype = C.c_double
            L.ref_bench_run_repeats.argtypes = [C.POINTER(CScene), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _libs[kind] = L
    return _libs[kind]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefWorld:
    """A vanilla b2World built from a Scene (KIND "ref"), or the C restatement (PortWorld, KIND "port")."""

    KIND = "ref"

    def __init__(self, scene: Scene, record_events=False):
        self.scene = scene
        self.L = lib(self.KIND)
        self.h = self.L.ref_world_create(C.byref(scene.c_scene()))
        if not self.h:
            raise RuntimeError(f"the {self.KIND} oracle cannot build this scene (joints need the live reference)")
        if record_events:
            self.L.ref_world_record_events(self.h, 1)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_world_destroy(self.h)
            self.h = None

    def step(self, dt, vel_iters, pos_iters):
        self.L.ref_world_step(self.h, dt, vel_iters, pos_iters)

    def command(self, body, op, wake=1, v=(0.0, 0.0), p=(0.0, 0.0)):
        """One between-step b2Body call (op = B2GPU_CMD_* of include/b2gpu.h)."""
        self.L.ref_world_body_command(self.h, body, op, wake, v[0], v[1], p[0], p[1])

    def query_aabb(self, lower, upper):
        """b2World::QueryAABB: fixture | child << 16 per reported proxy, in the reference's order (live reference only)."""
        lo, hi = np.asarray(lower, np.float32), np.asarray(upper, np.float32)
        out = np.zeros(4096, np.int32)
        n = self.L.ref_world_query_aabb(self.h, _p(lo), _p(hi), _p(out), out.shape[0])
        return out[:n].copy()

    def raycast_closest(self, p1, p2):
        """b2World::RayCast with the closest-hit callback: (fixture, child, fraction, point[2], normal[2]) or None."""
        a, b = np.asarray(p1, np.float32), np.asarray(p2, np.float32)
        oi, of = np.zeros(2, np.int32), np.zeros(5, np.float32)
        if not self.L.ref_world_raycast_closest(self.h, _p(a), _p(b), _p(oi), _p(of)):
            return None
        return int(oi[0]), int(oi[1]), of[0], of[1:3].copy(), of[3:5].copy()

    def create_body(self, body_def, fixture_defs, vertices):
        """b2World::CreateBody + CreateFixture after stepping (live reference only); returns the body index."""
        from box2dpp_b200.scene import FixtureDef
        fds = (FixtureDef * max(len(fixture_defs), 1))(*fixture_defs)
        v = np.ascontiguousarray(vertices if len(vertices) else [0.0, 0.0], np.float32)
        return self.L.ref_world_create_body(self.h, C.byref(body_def), fds, len(fixture_defs), _p(v))

    def set_filter(self, fixture, category_bits, mask_bits, group_index):
        """b2Fixture::SetFilterData (live reference only)."""
        self.L.ref_world_set_filter(self.h, fixture, category_bits, mask_bits, group_index)

    def set_density(self, fixture, density):
        """b2Fixture::SetDensity (live reference only); the body's mass follows at the next ResetMassData."""
        self.L.ref_world_set_density(self.h, fixture, density)

    def fixture_mass_data(self, fixture):
        """b2Fixture::GetMassData (live reference only): float32 [mass, I, center.x, center.y]."""
        out = np.zeros(4, np.float32)
        if self.L.ref_world_fixture_mass_data(self.h, fixture, _p(out)) != 0:
            raise IndexError(fixture)
        return out

    def create_joint(self, joint_def):
        """b2World::CreateJoint after stepping (live reference only); returns the joint id (creation order)."""
        return self.L.ref_world_create_joint(self.h, C.byref(joint_def))

    def create_fixture(self, body, fixture_def, vertices):
        """b2Body::CreateFixture on an existing body after stepping (live reference only); returns the fixture index."""
        v = np.ascontiguousarray(vertices if len(vertices) else [0.0, 0.0], np.float32)
        return self.L.ref_world_create_fixture(self.h, body, C.byref(fixture_def), _p(v))

    def bodies(self) -> np.ndarray:
        """[nb,16] float32: xf.p(2) q.s q.c | c(2) a | v(2) w | awake mass I localCenter(2) pad."""
        nb = self.L.ref_world_body_count(self.h)
        out = np.zeros((nb, 16), np.float32)
        self.L.ref_world_get_bodies(self.h, _p(out))
        return out

    def contacts(self):
        """(ints [n,8], floats [n,16]) in world-list order; see ref_harness.cpp."""
        n = self.L.ref_world_contact_count(self.h)
        ints = np.zeros((max(n, 1), 8), np.int32)
        fl = np.zeros((max(n, 1), 16), np.float32)
        self.L.ref_world_get_contacts(self.h, _p(ints), _p(fl), n)
        return ints[:n], fl[:n]

    def proxies(self):
        cap = self.L.ref_world_proxy_count(self.h) + 4096  # proxies of destroyed bodies keep their (zeroed) rows
        aabbs = np.zeros((cap, 4), np.float32)
        ids = np.zeros(cap, np.int32)
        n = min(self.L.ref_world_get_proxies(self.h, _p(aabbs), _p(ids), cap), cap)
        return aabbs[:n].copy(), ids[:n].copy()

    def events(self):
        cap = 1 << 16
        ints = np.zeros((cap, 6), np.int32)
        fl = np.zeros((cap, 4), np.float32)
        n = self.L.ref_world_get_events(self.h, _p(ints), _p(fl), cap)
        return ints[:n], fl[:n]


class PortWorld(RefWorld):
    KIND = "port"

    def islands(self):
        """(island count, body order over all islands, contact count) of the last step."""
        nb = self.L.ref_world_body_count(self.h)
        bodies = np.zeros(4 * (nb + 1), np.int32)
        counts = np.zeros(3, np.int32)
        n = self.L.ref_world_get_islands(self.h, _p(bodies), bodies.size, _p(counts))
        return int(counts[0]), bodies[:n].copy(), int(counts[2])


def polygon_build(fd: FixtureDef, verts: np.ndarray, kind="ref"):
    out = np.zeros(38, np.float32)
    v = np.ascontiguousarray(verts, np.float32)
    n = lib(kind).ref_polygon_build(C.byref(fd), _p(v), _p(out))
    return n, out


def _collide(fn, a, xa, b, xb, kind="ref"):
    ints = np.zeros(4, np.int32)
    fl = np.zeros(8, np.float32)
    a, xa, b, xb = (np.ascontiguousarray(t, np.float32) for t in (a, xa, b, xb))
    getattr(lib(kind), fn)(_p(a), _p(xa), _p(b), _p(xb), _p(ints), _p(fl))
    return ints, fl


def collide_polygons(pa, xa, pb, xb, kind="ref"):
    return _collide("ref_collide_polygons", pa, xa, pb, xb, kind)


def collide_circles(ca, xa, cb, xb, kind="ref"):
    return _collide("ref_collide_circles", ca, xa, cb, xb, kind)


def collide_polygon_circle(pa, xa, cb, xb, kind="ref"):
    return _collide("ref_collide_polygon_circle", pa, xa, cb, xb, kind)


def collide_edge_circle(ea, xa, cb, xb, kind="ref"):
    return _collide("ref_collide_edge_circle", ea, xa, cb, xb, kind)


def collide_edge_polygon(ea, xa, pb, xb, kind="ref"):
    return _collide("ref_collide_edge_polygon", ea, xa, pb, xb, kind)


def sincosf(x: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    lib().ref_sincosf(_p(x), _p(s), _p(c), x.size)
    return s, c


def bench_run(scenes, worlds_per_thread, threads, steps, warmup, dt, vel_iters, pos_iters):
    """All-core CPU baseline.  Returns (seconds, checksum)."""
    arr = scene_array(scenes)
    chk = C.c_uint64(0)
    sec = lib().ref_bench_run(arr, len(scenes), worlds_per_thread, threads, steps, warmup, dt, vel_iters,
                              pos_iters, C.byref(chk))
    return sec, chk.value


def bench_run_repeats(scenes, worlds_per_thread, threads, steps, warmup, dt, vel_iters, pos_iters, repeats=3):
    """All-core CPU baseline, the timed block of `steps` steps repeated: returns the list of seconds per block."""
    arr = scene_array(scenes)
    sec = (C.c_double * repeats)()
    lib().ref_bench_run_repeats(arr, len(scenes), worlds_per_thread, threads, steps, warmup, dt, vel_iters, pos_iters,
                                repeats, sec)
    return list(sec)


# ---- hashes of SURVEY.md Appendix D -------------------------------------------------------------
def fnv1a64(data: bytes) -> int:
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def state_hash(bodies: np.ndarray, skip_first=1) -> int:
    """FNV-1a over bits of x, y, angle, vx, vy, w per body in creation order, ground excluded."""
    b = bodies[skip_first:]
    cols = np.stack([b[:, 0], b[:, 1], b[:, 6], b[:, 7], b[:, 8], b[:, 9]], axis=1).astype("<f4")
    return fnv1a64(cols.tobytes())


def contact_order_hash(ints: np.ndarray, fixture_body: np.ndarray) -> int:
    """FNV-1a over (body index A, body index B, pointCount) as int32 per contact in world-list order."""
    rows = np.stack([fixture_body[ints[:, 0]], fixture_body[ints[:, 2]], ints[:, 5]], axis=1).astype("<i4")
    return fnv1a64(rows.tobytes())
