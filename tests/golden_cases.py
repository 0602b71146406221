"""Deterministic inputs of the golden fixtures (tests/golden/, written by tools/make_golden.py).

Everything here is integer / IEEE-float32 numpy arithmetic from fixed seeds, so the same inputs are
regenerated bit-for-bit on any box; only the ANSWERS are frozen in tests/golden/.
"""
import numpy as np

from box2dpp_b200 import scene as S

F32 = np.float32
NOCCD = S.WORLD_DEFAULT & ~S.WORLD_CONTINUOUS


# ---- shapes in the flat formats of oracle/ref_harness.cpp ---------------------------------------
def xf(px, py, angle):
    """b2Transform::Set(p, angle) with float32 sin/cos of numpy (inputs only; any unit-ish rotation is fine)."""
    a = F32(angle)
    return np.array([px, py, np.sin(a, dtype=F32), np.cos(a, dtype=F32)], F32)


def poly_from_verts(v):
    """35 floats: count, 8 verts, 8 normals, centroid.  Normals = normalised edge x 1 (b2PolygonShape.cpp:224-232)."""
    v = np.asarray(v, F32)
    n = len(v)
    out = np.zeros(35, F32)
    out[0] = n
    out[1:1 + 2 * n] = v.reshape(-1)
    e = np.roll(v, -1, axis=0) - v
    nrm = np.stack([e[:, 1], -e[:, 0]], axis=1).astype(F32)
    ln = np.sqrt((nrm[:, 0] * nrm[:, 0] + nrm[:, 1] * nrm[:, 1]).astype(F32), dtype=F32)
    nrm = (nrm * (F32(1.0) / ln)[:, None]).astype(F32)
    out[17:17 + 2 * n] = nrm.reshape(-1)
    out[33:35] = v.mean(axis=0, dtype=F32)
    return out


def box_poly(hx, hy):
    return poly_from_verts([[-hx, -hy], [hx, -hy], [hx, hy], [-hx, hy]])


def random_convex(rs, rmin=0.4, rmax=1.2):
    n = int(rs.randint(3, 9))
    ang = np.sort(rs.uniform(0, 2 * np.pi, n))
    # keep neighbouring vertices apart so the polygon is well-formed
    ang = (np.arange(n) + rs.uniform(0.1, 0.9, n)) * (2 * np.pi / n)
    r = rs.uniform(rmin, rmax)
    return poly_from_verts(np.stack([r * np.cos(ang), r * np.sin(ang)], axis=1))


def circle(rs):
    return np.array([rs.uniform(-0.2, 0.2), rs.uniform(-0.2, 0.2), rs.uniform(0.2, 1.0)], F32)


def edge(rs):
    x0 = rs.uniform(-3, -1)
    pts = np.cumsum(np.stack([rs.uniform(0.5, 2.0, 4), rs.uniform(-0.8, 0.8, 4)], axis=1), axis=0) + [x0, 0.0]
    e = np.zeros(10, F32)
    e[:8] = pts.reshape(-1)
    e[8] = rs.randint(0, 2)
    e[9] = rs.randint(0, 2)
    return e


def _pairs(seed, n, make_a, make_b, spread, a_is_edge=False):
    rs = np.random.RandomState(seed)
    cases = []
    for _ in range(n):
        a, b = make_a(rs), make_b(rs)
        xa = xf(rs.uniform(-0.5, 0.5), rs.uniform(-0.5, 0.5), rs.uniform(-3.2, 3.2))
        cx, cy = 0.0, 0.0
        if a_is_edge:  # aim shape B at a random point of the edge v1-v2 (in world space)
            t = rs.uniform(-0.2, 1.2)
            lx, ly = a[2] + t * (a[4] - a[2]), a[3] + t * (a[5] - a[3])
            cx = xa[3] * lx - xa[2] * ly + xa[0]
            cy = xa[2] * lx + xa[3] * ly + xa[1]
        xb = xf(cx + rs.uniform(-spread, spread), cy + rs.uniform(-spread, spread), rs.uniform(-3.2, 3.2))
        cases.append((a, xa, b, xb))
    return cases


def polygons_cases():
    # Regression/tests/helloworld.cpp:393-418 (Testb2CollidePolygonsTouching) and :358-391 geometry
    kat = [(box_poly(1, 1), xf(1, 0, 1.0), box_poly(1, 1), xf(-1, 0, 1.0)),
           (box_poly(1, 1), xf(1, 0, 0.0), box_poly(1, 1), xf(-1, 0, 0.0)),
           (box_poly(1, 1), xf(2, 0, 1.0), box_poly(1, 1), xf(-1, 0, 1.0)),
           # pyramid neighbours: unit boxes side by side / stacked (BASELINE config 2 geometry)
           (box_poly(0.5, 0.5), xf(0, 0, 0.0), box_poly(0.5, 0.5), xf(1.125, 0, 0.0)),
           (box_poly(0.5, 0.5), xf(0, 0, 0.0), box_poly(0.5, 0.5), xf(0.5625, 1.0, 0.0))]
    boxes = _pairs(11, 300, lambda rs: box_poly(rs.uniform(0.3, 1), rs.uniform(0.3, 1)),
                   lambda rs: box_poly(rs.uniform(0.3, 1), rs.uniform(0.3, 1)), 1.6)
    return kat + boxes + _pairs(12, 500, random_convex, random_convex, 2.0)


def circles_cases():
    # Regression/tests/helloworld.cpp:268-293 (Testb2CollideCirclesTouching, vanilla arm: xf1 at the origin)
    unit = np.array([0, 0, 1], F32)
    kat = [(unit, xf(0, 0, 1.0), unit, xf(-1, 0, 1.0)), (unit, xf(1, 0, 1.0), unit, xf(-1, 0, 1.0))]
    return kat + _pairs(13, 300, circle, circle, 1.5)


def polygon_circle_cases():
    return _pairs(14, 500, random_convex, circle, 1.8)


def edge_circle_cases():
    return _pairs(15, 500, edge, circle, 0.9, a_is_edge=True)


def edge_polygon_cases():
    return _pairs(16, 800, edge, random_convex, 1.0, a_is_edge=True)


NARROW_CASES = {
    "polygons": ("collide_polygons", polygons_cases),
    "circles": ("collide_circles", circles_cases),
    "polygon_circle": ("collide_polygon_circle", polygon_circle_cases),
    "edge_circle": ("collide_edge_circle", edge_circle_cases),
    "edge_polygon": ("collide_edge_polygon", edge_polygon_cases),
}


def polygon_cases():
    """(FixtureDef, vertices) for b2PolygonShape::SetAsBox / Set + ComputeMass."""
    rs = np.random.RandomState(21)
    cases = []
    for i in range(200):
        fd = S.FixtureDef()
        fd.shapeType = S.SHAPE_POLYGON
        fd.density = float(F32(rs.uniform(0.5, 5.0)))
        kind = i % 3
        verts = np.zeros((8, 2), F32)
        if kind == 0:
            fd.shapeKind = S.POLY_BOX
            fd.p[0], fd.p[1] = float(F32(rs.uniform(0.1, 3))), float(F32(rs.uniform(0.1, 3)))
        elif kind == 1:
            fd.shapeKind = S.POLY_BOX_OFFSET
            fd.p[0], fd.p[1] = float(F32(rs.uniform(0.1, 3))), float(F32(rs.uniform(0.1, 3)))
            fd.center[0], fd.center[1] = float(F32(rs.uniform(-2, 2))), float(F32(rs.uniform(-2, 2)))
            fd.angle = float(F32(rs.uniform(-3, 3)))
        else:
            fd.shapeKind = S.POLY_SET
            n = int(rs.randint(3, 9))
            pts = rs.uniform(-1.5, 1.5, (n, 2)).astype(F32)  # arbitrary cloud: exercises hull + welding
            if i % 9 == 2 and n > 3:
                pts[1] = pts[0] + F32(1e-4)  # welded duplicate (b2PolygonShape.cpp:136)
            verts[:n] = pts
            fd.firstVertex, fd.vertexCount = 0, n
        cases.append((fd, verts[:max(fd.vertexCount, 1)].copy()))
    return cases


def sincos_inputs():
    rs = np.random.RandomState(31)
    parts = [rs.uniform(-0.8, 0.8, 60000), rs.uniform(-8, 8, 60000), rs.uniform(-130, 130, 30000),
             rs.uniform(-1e6, 1e6, 20000), np.ldexp(rs.uniform(-1, 1, 20000), rs.randint(-40, 100, 20000)),
             [0.0, -0.0, np.pi / 4, np.pi / 2, np.pi, 1e-5, 2.44140625e-4, 120.0, 119.99999, 3.4e38, np.inf, np.nan]]
    x = np.concatenate([np.asarray(p, np.float64) for p in parts]).astype(F32)
    bits = rs.randint(0, 2 ** 32, 60000, dtype=np.uint64).astype(np.uint32).view(F32)  # raw bit patterns
    return np.concatenate([x, bits])


# ---- world scenes --------------------------------------------------------------------------------
WORLD_CASES = {
    # name: (scene factory, steps, velocity iterations, position iterations)
    "hello_world_noccd": (lambda: S.hello_world(NOCCD), 120, 6, 2),
    "hello_world_default": (lambda: S.hello_world(), 60, 6, 2),
    "pyramid20_noccd_sleep": (lambda: S.pyramid(20, flags=NOCCD), 600, 8, 3),
    "pyramid20_noccd_nosleep": (lambda: S.pyramid(20, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP), 600, 8, 3),
    "pyramid20_noccd_nosleep_62": (lambda: S.pyramid(20, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP), 600, 6, 2),
    "pyramid20_default": (lambda: S.pyramid(20), 600, 8, 3),
    "pyramid20_ccd_nosleep_62": (lambda: S.pyramid(20, flags=S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP), 600, 6, 2),
    "pyramid8_edge_ground": (lambda: S.pyramid(8, flags=NOCCD, ground="edge"), 250, 8, 3),
    "pyramid7_jitter3": (lambda: S.pyramid(7, flags=NOCCD, jitter_seed=3), 200, 8, 3),
    "mixed_seed0": (lambda: S.mixed(0, flags=NOCCD), 400, 8, 3),
    "mixed_seed5": (lambda: S.mixed(5, flags=NOCCD), 400, 8, 3),
    # continuous physics on (b2World::SolveTOI): the reference's default flags
    "mixed_seed2_default": (lambda: S.mixed(2), 300, 8, 3),
    "bullets_seed1_default": (lambda: S.bullets(1), 240, 8, 3),
    "bullets_seed3_default_nosleep": (lambda: S.bullets(3, flags=S.WORLD_DEFAULT & ~S.WORLD_ALLOW_SLEEP), 240, 6, 2),
    # sensor fixtures (b2Contact::Update's b2TestOverlap branch); live-reference only, like continuous physics
    "sensors_seed1_default": (lambda: S.sensors(1), 300, 8, 3),
    "sensors_seed2_default_noccd": (lambda: S.sensors(2, flags=NOCCD), 300, 6, 2),
    # joints in the island solver (revolute, prismatic, distance); live-reference only
    "jointed_seed1_default": (lambda: S.jointed(1), 400, 8, 3),
    "jointed_seed2_default_noccd": (lambda: S.jointed(2, flags=NOCCD), 400, 6, 2),
    "vehicles_seed1_default": (lambda: S.vehicles(1), 400, 8, 3),
    "vehicles_seed2_default_noccd": (lambda: S.vehicles(2, flags=NOCCD), 400, 6, 2),
    # b2Filter category / mask / group rules, linear and angular damping, gravityScale, fixedRotation (round 2)
    "filtered_seed1_noccd": (lambda: S.filtered(1, flags=NOCCD), 300, 8, 3),
    "filtered_seed4_noccd_nosleep_62": (lambda: S.filtered(4, flags=NOCCD & ~S.WORLD_ALLOW_SLEEP), 300, 6, 2),
    "filtered_seed2_default": (lambda: S.filtered(2), 300, 8, 3),
}
# scenes the C restatement can run (continuous physics off); the device path runs every case
NOCCD_CASES = [k for k in WORLD_CASES if "default" not in k and "_ccd_" not in k]

# SURVEY.md Appendix D checkpoints (vanilla, produced by the survey's independent probe):
# name -> {step: (contacts, touching, points, stateHash, contactOrderHash)}
APPENDIX_D = {
    "pyramid20_default": {0: (190, 0, 0, "8029c41d993d0540", "b2293f44fdd1c0eb"),
                          1: (None, None, None, "ba2eae883e3115ac", None),
                          59: (599, 398, 796, "6dc46e41b36d9f98", "8b2cb918d04239a7"),
                          599: (590, 400, 800, "da61302b8bd02db9", "9013bba0f35e0b31")},
    "pyramid20_noccd_sleep": {0: (190, 0, 0, "8029c41d993d0540", "b2293f44fdd1c0eb"),
                              59: (599, 398, 795, "951c38d230a541cb", "c490b1ad167046a4"),
                              599: (590, 400, 800, "046e71a23fe6c2c3", "9013bba0f35e0b31")},
    "pyramid20_noccd_nosleep": {59: (599, 398, 795, "951c38d230a541cb", "c490b1ad167046a4"),
                                599: (590, 400, 800, "7188226387729539", "9013bba0f35e0b31")},
    "pyramid20_noccd_nosleep_62": {59: (596, 397, 788, "ce0b6a6bbffaa049", "3dee0fc0bdfb5493"),
                                   599: (590, 401, 801, "66c33ff73f1ccf6f", "f11e7b5ad148d4c0")},
    "pyramid20_ccd_nosleep_62": {59: (604, 393, 783, "7bb7c20b0f899fe3", "a73fa00db93e2619"),
                                 599: (590, 401, 800, "23d01e2d0929de3a", "e0f478a55f95dba3")},
}
# HelloWorld with reference defaults, SURVEY.md Appendix D: step -> bits of y
APPENDIX_D_HELLO_Y = {0: 0x407fd27d, 29: 0x402d5555, 44: 0x3f900003, 45: 0x3f81ddd4, 59: 0x3f81ea67}


def fixture_bodies(scene):
    return np.array([f.body for f in scene.fixtures], np.int32)


_FNV_OFF, _FNV_PRIME = 0xCBF29CE484222325, 0x100000001B3


def fnv1a64(data: bytes) -> int:
    h = _FNV_OFF
    for b in data:
        h = ((h ^ b) * _FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
    return h


def state_hash(bodies, skip_first=1):
    """FNV-1a over bits of x, y, angle, vx, vy, w per body in creation order, ground excluded (Appendix D)."""
    b = bodies[skip_first:]
    cols = np.stack([b[:, 0], b[:, 1], b[:, 6], b[:, 7], b[:, 8], b[:, 9]], axis=1).astype("<f4")
    return fnv1a64(cols.tobytes())


def contact_order_hash(ints, fixture_body):
    """FNV-1a over (body A, body B, pointCount) int32 per contact in world-list order (Appendix D)."""
    if len(ints) == 0:
        return _FNV_OFF
    rows = np.stack([fixture_body[ints[:, 0]], fixture_body[ints[:, 2]], ints[:, 5]], axis=1).astype("<i4")
    return fnv1a64(rows.tobytes())
