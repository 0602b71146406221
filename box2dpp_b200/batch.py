"""ctypes binding of libb2gpu.so (include/b2gpu.h) for tests and bench.py.

``WorldBatch`` is the Python face of the C++ ``b2WorldBatch``: a batch of independent worlds built from
``Scene`` descriptions (the reference's CreateBody / CreateFixture calls) and stepped on one B200 by
``b2gpu_batch_step`` — the replacement of ``b2World::Step`` (Box2D/Dynamics/b2World.cpp:888-954).
There is no CPU fallback: construction raises if the CUDA library is missing or no device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scene import CScene, FixtureDef, Scene, SceneBatch, scene_array

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2GPU_LIB selects another build of the same library (diagnostic variants: tools/build_variant.sh)
LIB_PATH = os.environ.get("B2GPU_LIB") or os.path.join(_HERE, "libb2gpu.so")

# field ids — keep in sync with include/b2gpu.h
(F_GRAVITY, F_WORLD_FLAGS, F_INV_DT0, F_BODY_COUNT, F_PROXY_COUNT, F_CONTACT_COUNT, F_STATUS, F_EVENT_COUNT,
 F_ISLAND_COUNT, F_STEP_COUNT, F_BODY_XF, F_BODY_SWEEP, F_BODY_SWEEP0, F_BODY_VEL, F_BODY_FORCE, F_BODY_MASS,
 F_BODY_DAMP, F_BODY_FLAGS, F_FIXTURE_INFO, F_FIXTURE_SHAPE, F_FIXTURE_MAT, F_FIXTURE_FILTER, F_PROXY_FAT,
 F_PROXY_INFO, F_PROXY_MOVED, F_SHAPE_WORDS, F_CHAIN_VERTS, F_CONTACT_A, F_CONTACT_K, F_CONTACT_S, F_CONTACT_P0,
 F_CONTACT_P1, F_CONTACT_N, F_EVENT_A, F_EVENT_B, F_ISLAND_CONTACTS, F_ISLAND_CONTACT_END, F_JOINT_COUNT, F_JOINT_I,
 F_JOINT_F0, F_JOINT_F1, F_JOINT_F2, F_JOINT_IMPULSE, F_FIXTURE_COUNT, F_SHAPE_WORD_COUNT, F_CHAIN_VERT_COUNT,
 F_PROXY_FREE_COUNT, F_PROXY_FREE, F_NEW_FIXTURE, F_JOINT_SEQ, F_FIXTURE_MASS, F_COUNT) = range(52)

INT_FIELDS = {F_WORLD_FLAGS, F_BODY_COUNT, F_PROXY_COUNT, F_CONTACT_COUNT, F_STATUS, F_EVENT_COUNT, F_ISLAND_COUNT,
              F_STEP_COUNT, F_BODY_FLAGS, F_FIXTURE_INFO, F_FIXTURE_SHAPE, F_FIXTURE_FILTER, F_PROXY_INFO,
              F_PROXY_MOVED, F_CONTACT_A, F_CONTACT_K, F_CONTACT_S, F_EVENT_A, F_ISLAND_CONTACTS,
              F_ISLAND_CONTACT_END, F_JOINT_COUNT, F_JOINT_I, F_FIXTURE_COUNT, F_SHAPE_WORD_COUNT, F_CHAIN_VERT_COUNT,
              F_PROXY_FREE_COUNT, F_PROXY_FREE, F_NEW_FIXTURE, F_JOINT_SEQ}

EVENT_BEGIN, EVENT_END, EVENT_PRESOLVE, EVENT_POSTSOLVE = 0, 1, 2, 3
EVENT_ALL = 15

EXPORTED_SYMBOLS = [
    "b2gpu_batch_create", "b2gpu_batch_destroy", "b2gpu_batch_upload_field", "b2gpu_batch_download_field",
    "b2gpu_batch_field_shape", "b2gpu_batch_step", "b2gpu_batch_step_n", "b2gpu_batch_sync",
    "b2gpu_batch_apply_forces", "b2gpu_batch_read_transforms", "b2gpu_batch_step_io", "b2gpu_batch_io_select", "b2gpu_batch_last_step_ms",
    "b2gpu_batch_profile_steps", "b2gpu_batch_profile_kernels", "b2gpu_batch_kernel_name", "b2gpu_batch_launch_count", "b2gpu_last_error", "b2gpu_scene_capacities",
    "b2gpu_batch_build_worlds", "b2gpu_host_polygon_build", "b2gpu_host_fixture_mass_data", "b2gpu_host_sincosf", "b2gpu_sincosf",
    "b2gpu_collide_pairs", "b2gpu_batch_create_bodies", "b2gpu_batch_create_fixtures", "b2gpu_batch_create_joints",
    "b2gpu_batch_joint_records", "b2gpu_batch_spawn_records", "b2gpu_batch_body_commands",
]

COLLIDE_POLYGONS, COLLIDE_CIRCLES, COLLIDE_POLYGON_CIRCLE, COLLIDE_EDGE_CIRCLE, COLLIDE_EDGE_POLYGON = range(5)


class BatchDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "device", "numWorlds", "maxBodies", "maxFixtures", "maxProxies", "maxContacts", "maxShapeWords",
        "maxChainVerts", "maxEvents", "eventMask", "blockSize", "reserved", "maxJoints")]


class B2GpuError(RuntimeError):
    pass


_lib = None


# b2gpu_body_command (include/b2gpu.h)
BODY_COMMAND_DTYPE = np.dtype([("world", "<i4"), ("body", "<i4"), ("op", "<i4"), ("wake", "<i4"),
                               ("v", "<f4", 2), ("p", "<f4", 2)])
(CMD_APPLY_FORCE, CMD_APPLY_FORCE_TO_CENTER, CMD_APPLY_TORQUE, CMD_APPLY_LINEAR_IMPULSE, CMD_APPLY_ANGULAR_IMPULSE,
 CMD_SET_LINEAR_VELOCITY, CMD_SET_ANGULAR_VELOCITY, CMD_SET_TRANSFORM, CMD_SET_AWAKE, CMD_DESTROY_BODY,
 CMD_MOUSE_SET_TARGET, CMD_SET_FILTER, CMD_DESTROY_JOINT, CMD_SET_TYPE, CMD_SET_ACTIVE, CMD_SET_MASS_DATA,
 CMD_RESET_MASS_DATA, CMD_DESTROY_FIXTURE, CMD_SET_DAMPING, CMD_SET_GRAVITY_SCALE, CMD_SET_BULLET, CMD_SET_SLEEPING_ALLOWED,
 CMD_SET_FIXED_ROTATION, CMD_SET_SENSOR, CMD_SET_MATERIAL, CMD_SET_DENSITY) = range(26)


def put_filter(commands: np.ndarray, i: int, category_bits=1, mask_bits=0xffff, group_index=0):
    """Fills the v field of the CMD_SET_FILTER command commands[i]: two int32 BIT PATTERNS in the float slots (written through an
    integer view: a float assignment could quieten a NaN pattern)."""
    bits = commands["v"].view(np.uint32)
    bits[i, 0] = (category_bits & 0xffff) | ((mask_bits & 0xffff) << 16)
    bits[i, 1] = np.array([group_index], np.int32).view(np.uint32)[0]



AABB_QUERY_DTYPE = np.dtype([("world", "<i4"), ("lower", "<f4", 2), ("upper", "<f4", 2)])       # b2gpu_aabb_query
RAY_DTYPE = np.dtype([("world", "<i4"), ("p1", "<f4", 2), ("p2", "<f4", 2), ("maxFraction", "<f4")])  # b2gpu_ray
RAY_HIT_DTYPE = np.dtype([("fixture", "<i4"), ("child", "<i4"), ("fraction", "<f4"), ("point", "<f4", 2),
                          ("normal", "<f4", 2)])                                                      # b2gpu_ray_hit


EVENT_DTYPE = np.dtype([("type", "<i4"), ("fixtureA", "<i4"), ("childA", "<i4"), ("fixtureB", "<i4"), ("childB", "<i4"),
                        ("count", "<i4"), ("normalImpulses", "<f4", 2), ("tangentImpulses", "<f4", 2)])  # b2gpu_event


def lib():
    """Load libb2gpu.so.  Raises (never falls back) when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2GpuError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                             "box2dpp_b200 has no CPU path")
        L = C.CDLL(LIB_PATH)
        vp, i32 = C.c_void_p, C.c_int32
        L.b2gpu_batch_create.argtypes = [C.POINTER(BatchDesc), C.POINTER(vp)]
        L.b2gpu_batch_destroy.argtypes = [vp]
        L.b2gpu_batch_upload_field.argtypes = [vp, i32, vp, i32, i32]
        L.b2gpu_batch_download_field.argtypes = [vp, i32, vp, i32, i32]
        L.b2gpu_batch_field_shape.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
        L.b2gpu_batch_step.argtypes = [vp, C.c_float, i32, i32]
        L.b2gpu_batch_step_n.argtypes = [vp, C.c_float, i32, i32, i32]
        L.b2gpu_batch_sync.argtypes = [vp]
        L.b2gpu_batch_apply_forces.argtypes = [vp, vp]
        L.b2gpu_batch_read_transforms.argtypes = [vp, vp]
        L.b2gpu_batch_step_io.argtypes = [vp, vp, C.c_float, i32, i32, vp]
        if hasattr(L, "b2gpu_batch_io_select"):  # absent from older diagnostic builds (B2GPU_LIB)
            L.b2gpu_batch_io_select.argtypes = [vp, vp, i32, vp, i32, i32]
        L.b2gpu_batch_body_commands.argtypes = [vp, vp, C.c_int64]
        L.b2gpu_batch_read_body_state.argtypes = [vp, vp]
        L.b2gpu_batch_query_aabb.argtypes = [vp, vp, C.c_int64, i32, vp, vp]
        L.b2gpu_batch_raycast_closest.argtypes = [vp, vp, C.c_int64, vp]
        L.b2gpu_batch_read_events.argtypes = [vp, vp, vp, C.c_int64]
        L.b2gpu_batch_snapshot_size.argtypes = [vp]
        L.b2gpu_batch_snapshot_size.restype = C.c_int64
        L.b2gpu_batch_snapshot_save.argtypes = [vp, vp, C.c_int64]
        L.b2gpu_batch_snapshot_load.argtypes = [vp, vp, C.c_int64]
        L.b2gpu_batch_create_bodies.argtypes = [vp, vp, i32, vp, vp, vp]
        L.b2gpu_batch_create_fixtures.argtypes = [vp, vp, i32, vp, vp, vp]
        L.b2gpu_batch_create_joints.argtypes = [vp, vp, vp, i32, vp]
        L.b2gpu_batch_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.b2gpu_batch_profile_steps.argtypes = [vp, C.c_float, i32, i32, i32, C.POINTER(C.c_float)]
        L.b2gpu_batch_profile_kernels.argtypes = [vp, C.c_float, i32, i32, i32, C.POINTER(C.c_float), i32, C.POINTER(i32)]
        L.b2gpu_batch_kernel_name.argtypes = [vp, i32]
        L.b2gpu_batch_kernel_name.restype = C.c_char_p
        L.b2gpu_batch_launch_count.argtypes = [vp]
        L.b2gpu_batch_launch_count.restype = C.c_int64
        L.b2gpu_last_error.restype = C.c_char_p
        L.b2gpu_scene_capacities.argtypes = [C.POINTER(CScene), i32, C.POINTER(BatchDesc)]
        L.b2gpu_batch_build_worlds.argtypes = [vp, C.POINTER(CScene), i32, i32, i32]
        L.b2gpu_host_polygon_build.argtypes = [C.POINTER(FixtureDef), vp, vp]
        L.b2gpu_host_fixture_mass_data.argtypes = [C.POINTER(FixtureDef), vp, i32, vp]
        L.b2gpu_host_sincosf.argtypes = [vp, vp, vp, C.c_int64]
        L.b2gpu_host_sincosf_wide.argtypes = [vp, vp, vp, C.c_int64]
        L.b2gpu_sincosf.argtypes = [i32, vp, vp, vp, C.c_int64]
        L.b2gpu_collide_pairs.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise B2GpuError(f"b2gpu error {rc}: {lib().b2gpu_last_error().decode()}")


def host_polygon_build(fd: FixtureDef, verts):
    v = np.ascontiguousarray(verts, np.float32)
    out = np.zeros(38, np.float32)
    n = lib().b2gpu_host_polygon_build(C.byref(fd), v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return n, out


def host_fixture_mass_data(fd: FixtureDef, verts, density=None):
    """b2gpu_host_fixture_mass_data: float32 [mass, I, center.x, center.y] of one fixture description (fd.firstVertex indexes
    `verts`, the scene's vertex list) at `density` (default: fd.density) — the payload of CMD_SET_DENSITY."""
    v = np.ascontiguousarray(verts if len(verts) else [0.0, 0.0], np.float32).reshape(-1)
    one = FixtureDef.from_buffer_copy(fd)
    if density is not None:
        one.density = density
    out = np.zeros(4, np.float32)
    _check(lib().b2gpu_host_fixture_mass_data(C.byref(one), v.ctypes.data_as(C.c_void_p), len(v) // 2, out.ctypes.data_as(C.c_void_p)))
    return out


def collide_pairs(kind, shape_a, xf_a, shape_b, xf_b, device=0):
    """b2gpu_collide_pairs: the device narrow phase on n independent shape pairs (flat formats of include/b2gpu.h)."""
    a, xa, b, xb = (np.ascontiguousarray(t, np.float32) for t in (shape_a, xf_a, shape_b, xf_b))
    n = xa.shape[0]
    ints = np.zeros((n, 4), np.int32)
    fl = np.zeros((n, 8), np.float32)
    p = lambda t: t.ctypes.data_as(C.c_void_p)  # noqa: E731
    _check(lib().b2gpu_collide_pairs(device, kind, n, p(a), p(xa), p(b), p(xb), p(ints), p(fl)))
    return ints, fl


class WorldBatch:
    """A batch of independent worlds on one GPU."""

    def __init__(self, scenes, replicate=1, device=0, max_contacts=0, max_events=0, event_mask=0, block_size=0,
                 generic_solver=False, always_rebuild_islands=False, binned_collide=None, headroom=None, joint_headroom=0):
        if isinstance(scenes, Scene):
            scenes = [scenes]
        if isinstance(scenes, SceneBatch):
            self.scenes = scenes
            arr, n_scenes = scenes.c_array, scenes.count
        else:
            self.scenes = list(scenes)
            arr, n_scenes = scene_array(self.scenes), len(self.scenes)
        self.num_worlds = n_scenes * replicate
        L = lib()
        d = BatchDesc()
        d.device, d.numWorlds = device, self.num_worlds
        d.maxContacts, d.maxEvents, d.eventMask, d.blockSize = max_contacts, max_events, event_mask, block_size
        d.reserved = (1 if generic_solver else 0) | (2 if always_rebuild_islands else 0)
        _check(L.b2gpu_scene_capacities(arr, n_scenes, C.byref(d)))
        if headroom:  # room for bodies created after the first step: (bodies, fixtures, proxies, shape words, chain vertices)
            d.maxBodies += headroom[0]; d.maxFixtures += headroom[1]; d.maxProxies += headroom[2]
            d.maxShapeWords += headroom[3]; d.maxChainVerts += headroom[4]
            d.maxContacts += 4 * headroom[2]
            if d.eventMask:
                d.maxEvents = max(d.maxEvents, 3 * d.maxContacts)
        d.maxJoints += joint_headroom  # room for b2World::CreateJoint on the running worlds
        if binned_collide is not None:  # default: decided by b2gpu_scene_capacities (worlds that mix shape types)
            d.reserved = (d.reserved | 4) if binned_collide else (d.reserved & ~4)
        self.desc = d
        self.h = C.c_void_p()
        _check(L.b2gpu_batch_create(C.byref(d), C.byref(self.h)))
        _check(L.b2gpu_batch_build_worlds(self.h, arr, n_scenes, 0, replicate))
        self.max_bodies = d.maxBodies
        self._pinned = {}
        self.io_force_count = self.io_pose_count = d.maxBodies  # b2gpu_batch_io_select
        self.io_pose_floats = 4

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().b2gpu_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- b2World::Step ----
    def step(self, dt, vel_iters, pos_iters):
        _check(lib().b2gpu_batch_step(self.h, dt, vel_iters, pos_iters))

    def step_n(self, dt, vel_iters, pos_iters, steps):
        _check(lib().b2gpu_batch_step_n(self.h, dt, vel_iters, pos_iters, steps))

    def sync(self):
        _check(lib().b2gpu_batch_sync(self.h))

    def last_step_ms(self) -> float:
        ms = C.c_float(0)
        _check(lib().b2gpu_batch_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def profile_steps(self, dt, vel_iters, pos_iters, steps):
        ms = (C.c_float * 4)()
        _check(lib().b2gpu_batch_profile_steps(self.h, dt, vel_iters, pos_iters, steps, ms))
        return dict(zip(("collide", "islands", "solve", "sync_pairs"), [float(x) for x in ms]))

    def profile_kernels(self, dt, vel_iters, pos_iters, steps):
        """{kernel name: total ms over `steps` steps}, in launch order (CUDA events around every kernel)."""
        ms = (C.c_float * 16)()
        n = C.c_int32(0)
        _check(lib().b2gpu_batch_profile_kernels(self.h, dt, vel_iters, pos_iters, steps, ms, 16, C.byref(n)))
        return {lib().b2gpu_batch_kernel_name(self.h, i).decode(): float(ms[i]) for i in range(n.value)}

    def launch_count(self) -> int:
        return int(lib().b2gpu_batch_launch_count(self.h))

    # ---- fields ----
    def field_shape(self, fid):
        cap, comps = C.c_int32(), C.c_int32()
        _check(lib().b2gpu_batch_field_shape(self.h, fid, C.byref(cap), C.byref(comps)))
        return cap.value, comps.value

    def field(self, fid, first=0, count=None) -> np.ndarray:
        count = self.num_worlds - first if count is None else count
        cap, comps = self.field_shape(fid)
        out = np.zeros((count, cap, comps), np.int32 if fid in INT_FIELDS else np.float32)
        _check(lib().b2gpu_batch_download_field(self.h, fid, out.ctypes.data_as(C.c_void_p), first, count))
        return out

    def set_field(self, fid, values, first=0):
        cap, comps = self.field_shape(fid)
        v = np.ascontiguousarray(values, np.int32 if fid in INT_FIELDS else np.float32).reshape(-1, cap, comps)
        _check(lib().b2gpu_batch_upload_field(self.h, fid, v.ctypes.data_as(C.c_void_p), first, v.shape[0]))

    # ---- per-step I/O ----
    def io_select(self, force_bodies=None, pose_bodies=None, xya=False):
        """b2gpu_batch_io_select: the bodies (index lists, the same in every world; None = every body slot, [] = none)
        whose forces go in / whose poses come out, and the pose format (xya: p.x p.y angle instead of p.x p.y q.s q.c)."""
        def arg(ids):
            if ids is None:
                return None, -1
            a = np.ascontiguousarray(ids, np.int32)
            return a, a.shape[0]
        fa, fn = arg(force_bodies)
        pa, pn = arg(pose_bodies)
        _check(lib().b2gpu_batch_io_select(self.h, fa.ctypes.data_as(C.c_void_p) if fa is not None else None, fn,
                                           pa.ctypes.data_as(C.c_void_p) if pa is not None else None, pn, 1 if xya else 0))
        self.io_force_count = self.max_bodies if fn < 0 else fn
        self.io_pose_count = self.max_bodies if pn < 0 else pn
        self.io_pose_floats = 3 if xya else 4

    def apply_forces(self, forces: np.ndarray):
        f = np.ascontiguousarray(forces, np.float32)
        assert f.size == self.num_worlds * self.io_force_count * 3
        _check(lib().b2gpu_batch_apply_forces(self.h, f.ctypes.data_as(C.c_void_p)))

    def step_io(self, forces: np.ndarray | None, dt: float, vel_iters: int, pos_iters: int,
                out: np.ndarray | None = None) -> np.ndarray:
        """b2gpu_batch_step_io: apply_forces + step + read_transforms in one call, the host copies of each world
        group overlapped with the other groups' kernels.  Returns the transforms [W, bodies, 4] (or what io_select chose)."""
        fptr = None
        if forces is not None:
            f = np.ascontiguousarray(forces, np.float32)
            assert f.size == self.num_worlds * self.io_force_count * 3
            fptr = f.ctypes.data_as(C.c_void_p)
        if out is None:
            out = np.empty((self.num_worlds, self.io_pose_count, self.io_pose_floats), np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        assert out.size == self.num_worlds * self.io_pose_count * self.io_pose_floats
        _check(lib().b2gpu_batch_step_io(self.h, fptr, dt, vel_iters, pos_iters, out.ctypes.data_as(C.c_void_p)))
        return out

    def body_commands(self, commands: np.ndarray):
        """b2gpu_batch_body_commands: `commands` is a BODY_COMMAND_DTYPE array (any world order)."""
        cmds = np.ascontiguousarray(commands, BODY_COMMAND_DTYPE)
        _check(lib().b2gpu_batch_body_commands(self.h, cmds.ctypes.data_as(C.c_void_p), cmds.shape[0]))

    def read_body_state(self, out: np.ndarray | None = None) -> np.ndarray:
        """[W, maxBodies, 8] float32: p.x p.y angle awake | v.x v.y omega sleepTime."""
        if out is None:
            out = np.empty((self.num_worlds, self.max_bodies, 8), np.float32)
        _check(lib().b2gpu_batch_read_body_state(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def query_aabb(self, queries: np.ndarray, max_hits=64):
        """b2gpu_batch_query_aabb -> (counts [n], hits [n, max_hits] fixture | child << 16, ascending proxy id)."""
        q = np.ascontiguousarray(queries, AABB_QUERY_DTYPE)
        counts = np.zeros(q.shape[0], np.int32)
        hits = np.zeros((q.shape[0], max_hits), np.int32)
        p = lambda t: t.ctypes.data_as(C.c_void_p)  # noqa: E731
        _check(lib().b2gpu_batch_query_aabb(self.h, p(q), q.shape[0], max_hits, p(counts), p(hits)))
        return counts, hits

    def raycast_closest(self, rays: np.ndarray) -> np.ndarray:
        """b2gpu_batch_raycast_closest -> RAY_HIT_DTYPE array (fixture = -1: no hit)."""
        r = np.ascontiguousarray(rays, RAY_DTYPE)
        out = np.zeros(r.shape[0], RAY_HIT_DTYPE)
        _check(lib().b2gpu_batch_raycast_closest(self.h, r.ctypes.data_as(C.c_void_p), r.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def read_transforms(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.num_worlds, self.io_pose_count, self.io_pose_floats), np.float32)
        assert out.size == self.num_worlds * self.io_pose_count * self.io_pose_floats
        _check(lib().b2gpu_batch_read_transforms(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    # ---- dumps in the oracle's formats (oracle/ref_harness.cpp) for parity tests ----
    def bodies(self, world) -> np.ndarray:
        nb = int(self.field(F_BODY_COUNT, world, 1)[0, 0, 0])
        xf = self.field(F_BODY_XF, world, 1)[0, :nb]
        sw = self.field(F_BODY_SWEEP, world, 1)[0, :nb]
        ve = self.field(F_BODY_VEL, world, 1)[0, :nb]
        ms = self.field(F_BODY_MASS, world, 1)[0, :nb]
        fl = self.field(F_BODY_FLAGS, world, 1)[0, :nb, 0]
        out = np.zeros((nb, 16), np.float32)
        out[:, 0:4] = xf
        out[:, 4:7] = sw[:, 0:3]
        out[:, 7:10] = ve[:, 0:3]
        out[:, 10] = ((fl & 8) != 0).astype(np.float32)
        with np.errstate(divide="ignore"):
            out[:, 11] = np.where(ms[:, 0] > 0, np.float32(1.0) / ms[:, 0], 0.0)
        out[:, 13:15] = ms[:, 2:4]
        dead = (fl & (1 << 10)) != 0  # destroyed bodies: a row of zeros with the marker in the last column (oracle format)
        out[dead] = 0.0
        out[dead, 15] = 1.0
        return out

    def contacts(self, world):
        n = int(self.field(F_CONTACT_COUNT, world, 1)[0, 0, 0])
        A = self.field(F_CONTACT_A, world, 1)[0, :n][::-1]
        K = self.field(F_CONTACT_K, world, 1)[0, :n][::-1]
        P0 = self.field(F_CONTACT_P0, world, 1)[0, :n][::-1]
        P1 = self.field(F_CONTACT_P1, world, 1)[0, :n][::-1]
        N = self.field(F_CONTACT_N, world, 1)[0, :n][::-1]
        pinfo = self.field(F_PROXY_INFO, world, 1)[0, :, 0]
        ints = np.zeros((n, 8), np.int32)
        fl = np.zeros((n, 16), np.float32)
        pa, pb = A[:, 0] & 0xFFFF, (A[:, 0] >> 16) & 0xFFFF
        ints[:, 0], ints[:, 1] = pinfo[pa] & 0xFFFF, (pinfo[pa] >> 16) & 0xFFFF
        ints[:, 2], ints[:, 3] = pinfo[pb] & 0xFFFF, (pinfo[pb] >> 16) & 0xFFFF
        flags = A[:, 2]
        ints[:, 4] = ((flags & 2) != 0) * 1 + ((flags & 4) != 0) * 2
        pc = (A[:, 3] >> 8) & 0xFF
        ints[:, 5] = pc
        ints[:, 6] = np.where(pc > 0, K[:, 0], 0)
        ints[:, 7] = np.where(pc > 1, K[:, 1], 0)
        has = pc > 0
        fl[:, 0:4] = np.where(has[:, None], N, 0)
        fl[:, 4] = np.where(has, (A[:, 3] & 0xFF).astype(np.float32), 0)
        fl[:, 5:9] = np.where(has[:, None], P0, 0)
        fl[:, 9:13] = np.where((pc > 1)[:, None], P1, 0)
        fl[:, 13] = K[:, 2].view(np.float32)
        fl[:, 14] = K[:, 3].view(np.float32)
        return ints, fl

    def proxies(self, world):
        n = int(self.field(F_PROXY_COUNT, world, 1)[0, 0, 0])
        fat = self.field(F_PROXY_FAT, world, 1)[0, :n].copy()
        fat[fat[:, 0] > fat[:, 2]] = 0.0  # proxies of destroyed bodies carry an empty box: zeros in the oracle format
        return fat

    def events(self, world):
        n = int(self.field(F_EVENT_COUNT, world, 1)[0, 0, 0])
        A = self.field(F_EVENT_A, world, 1)[0, :n]
        B = self.field(F_EVENT_B, world, 1)[0, :n]
        ints = np.zeros((n, 6), np.int32)
        ints[:, 0] = A[:, 0]
        ints[:, 1], ints[:, 2] = A[:, 1] & 0xFFFF, (A[:, 1] >> 16) & 0xFFFF
        ints[:, 3], ints[:, 4] = A[:, 2] & 0xFFFF, (A[:, 2] >> 16) & 0xFFFF
        ints[:, 5] = A[:, 3]
        return ints, B.copy()

    def read_events(self):
        """b2gpu_batch_read_events -> (offsets [W + 1] int64, events EVENT_DTYPE[offsets[-1]]) of the last step."""
        offsets = np.zeros(self.num_worlds + 1, np.int64)
        _check(lib().b2gpu_batch_read_events(self.h, offsets.ctypes.data_as(C.c_void_p), None, 0))
        events = np.zeros(int(offsets[-1]), EVENT_DTYPE)
        if events.shape[0]:
            _check(lib().b2gpu_batch_read_events(self.h, offsets.ctypes.data_as(C.c_void_p), events.ctypes.data_as(C.c_void_p),
                                                 events.shape[0]))
        return offsets, events

    def create_bodies(self, spawns):
        """b2gpu_batch_create_bodies: spawns = [(world, BodyDef, [FixtureDef, ...], vertices xy list)], vertex indices of a
        spawn's fixtures are relative to its own vertex list.  Returns the new body indices (-1: arenas full)."""
        from .scene import BodyDef, FixtureDef

        class Spawn(C.Structure):
            _fields_ = [("world", C.c_int32), ("firstFixture", C.c_int32), ("fixtureCount", C.c_int32), ("body", BodyDef)]

        n = len(spawns)
        arr = (Spawn * max(n, 1))()
        fds, verts = [], []
        for i, (world, bd, fixtures, v) in enumerate(spawns):
            arr[i].world, arr[i].firstFixture, arr[i].fixtureCount, arr[i].body = world, len(fds), len(fixtures), bd
            base = len(verts) // 2
            for f in fixtures:
                g = FixtureDef.from_buffer_copy(bytes(f))
                g.firstVertex += base
                fds.append(g)
            verts.extend(float(x) for x in v)
        cf = (FixtureDef * max(len(fds), 1))(*fds)
        cv = np.ascontiguousarray(verts if verts else [0.0, 0.0], np.float32)
        out = np.zeros(max(n, 1), np.int32)
        _check(lib().b2gpu_batch_create_bodies(self.h, arr, n, cf, cv.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out[:n]

    def create_fixtures(self, attach):
        """b2gpu_batch_create_fixtures: attach = [(world, body, [FixtureDef, ...], vertices xy list)] — b2Body::CreateFixture on
        bodies that exist.  Returns the index of each entry's first new fixture (-1: arenas full / no such body)."""
        from .scene import FixtureDef

        class Attach(C.Structure):
            _fields_ = [("world", C.c_int32), ("body", C.c_int32), ("firstFixture", C.c_int32), ("fixtureCount", C.c_int32)]

        n = len(attach)
        arr = (Attach * max(n, 1))()
        fds, verts = [], []
        for i, (world, body, fixtures, v) in enumerate(attach):
            arr[i].world, arr[i].body, arr[i].firstFixture, arr[i].fixtureCount = world, body, len(fds), len(fixtures)
            base = len(verts) // 2
            for f in fixtures:
                g = FixtureDef.from_buffer_copy(bytes(f))
                g.firstVertex += base
                fds.append(g)
            verts.extend(float(x) for x in v)
        cf = (FixtureDef * max(len(fds), 1))(*fds)
        cv = np.ascontiguousarray(verts if verts else [0.0, 0.0], np.float32)
        out = np.zeros(max(n, 1), np.int32)
        _check(lib().b2gpu_batch_create_fixtures(self.h, arr, n, cf, cv.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out[:n]

    def create_joints(self, joints):
        """b2gpu_batch_create_joints: joints = [(world, JointDef)] — b2World::CreateJoint on running worlds.  Returns joint ids."""
        from .scene import JointDef

        n = len(joints)
        worlds = np.ascontiguousarray([w for w, _ in joints] or [0], np.int32)
        defs = (JointDef * max(n, 1))(*[JointDef.from_buffer_copy(bytes(d)) for _, d in joints])
        out = np.zeros(max(n, 1), np.int32)
        _check(lib().b2gpu_batch_create_joints(self.h, worlds.ctypes.data_as(C.c_void_p), defs, n, out.ctypes.data_as(C.c_void_p)))
        return out[:n]

    def snapshot(self) -> np.ndarray:
        """b2gpu_batch_snapshot_save: every persistent field of every world as one blob."""
        blob = np.zeros(int(lib().b2gpu_batch_snapshot_size(self.h)), np.uint8)
        _check(lib().b2gpu_batch_snapshot_save(self.h, blob.ctypes.data_as(C.c_void_p), blob.shape[0]))
        return blob

    def restore(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, np.uint8)
        _check(lib().b2gpu_batch_snapshot_load(self.h, blob.ctypes.data_as(C.c_void_p), blob.shape[0]))

    def status(self):
        return self.field(F_STATUS)[:, 0, 0]
