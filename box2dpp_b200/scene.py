"""Scene descriptions (ctypes mirror of include/b2g_scene.h) and the synthetic scenes of BASELINE.json.

A ``Scene`` records the arguments of the reference's scene-construction calls
(``b2World::CreateBody`` Box2D/Dynamics/b2World.cpp:98-122, ``b2Body::CreateFixture``
Box2D/Dynamics/b2Body.cpp:168-214) in call order.  The same description feeds the product
(``box2dpp_b200.batch.WorldBatch``) and the CPU oracles under ``oracle/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

F32 = np.float32

STATIC_BODY, KINEMATIC_BODY, DYNAMIC_BODY = 0, 1, 2
SHAPE_CIRCLE, SHAPE_EDGE, SHAPE_POLYGON, SHAPE_CHAIN = 0, 1, 2, 3
POLY_BOX, POLY_BOX_OFFSET, POLY_SET = 0, 1, 2
CHAIN_OPEN, CHAIN_LOOP = 0, 1
BD_ALLOW_SLEEP, BD_AWAKE, BD_FIXED_ROTATION, BD_BULLET, BD_ACTIVE = 1, 2, 4, 8, 16
BD_DEFAULT = BD_ALLOW_SLEEP | BD_AWAKE | BD_ACTIVE
WORLD_ALLOW_SLEEP, WORLD_WARM_STARTING, WORLD_CONTINUOUS, WORLD_CLEAR_FORCES = 1, 2, 4, 8
WORLD_DEFAULT = 15


class BodyDef(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("flags", C.c_int32),
        ("position", C.c_float * 2),
        ("angle", C.c_float),
        ("linearVelocity", C.c_float * 2),
        ("angularVelocity", C.c_float),
        ("linearDamping", C.c_float),
        ("angularDamping", C.c_float),
        ("gravityScale", C.c_float),
    ]


class FixtureDef(C.Structure):
    _fields_ = [
        ("body", C.c_int32),
        ("shapeType", C.c_int32),
        ("shapeKind", C.c_int32),
        ("isSensor", C.c_int32),
        ("friction", C.c_float),
        ("restitution", C.c_float),
        ("density", C.c_float),
        ("categoryBits", C.c_uint16),
        ("maskBits", C.c_uint16),
        ("groupIndex", C.c_int16),
        ("pad0", C.c_int16),
        ("radius", C.c_float),
        ("p", C.c_float * 2),
        ("center", C.c_float * 2),
        ("angle", C.c_float),
        ("firstVertex", C.c_int32),
        ("vertexCount", C.c_int32),
        ("hasVertex0", C.c_int32),
        ("hasVertex3", C.c_int32),
        ("vertex0", C.c_float * 2),
        ("vertex3", C.c_float * 2),
    ]


JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_DISTANCE, JOINT_WHEEL, JOINT_WELD, JOINT_FRICTION, JOINT_ROPE, JOINT_MOTOR = 1, 2, 3, 7, 8, 9, 10, 11
JOINT_PULLEY, JOINT_MOUSE, JOINT_GEAR = 4, 5, 6


class JointDef(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("bodyA", C.c_int32),
        ("bodyB", C.c_int32),
        ("collideConnected", C.c_int32),
        ("localAnchorA", C.c_float * 2),
        ("localAnchorB", C.c_float * 2),
        ("localAxisA", C.c_float * 2),
        ("referenceAngle", C.c_float),
        ("enableLimit", C.c_int32),
        ("enableMotor", C.c_int32),
        ("lower", C.c_float),
        ("upper", C.c_float),
        ("maxMotor", C.c_float),
        ("motorSpeed", C.c_float),
        ("length", C.c_float),
        ("frequencyHz", C.c_float),
        ("dampingRatio", C.c_float),
    ]


class CScene(C.Structure):
    _fields_ = [
        ("gravity", C.c_float * 2),
        ("worldFlags", C.c_int32),
        ("bodyCount", C.c_int32),
        ("fixtureCount", C.c_int32),
        ("vertexCount", C.c_int32),
        ("bodies", C.POINTER(BodyDef)),
        ("fixtures", C.POINTER(FixtureDef)),
        ("vertices", C.POINTER(C.c_float)),
        ("joints", C.POINTER(JointDef)),
        ("jointCount", C.c_int32),
        ("pad0", C.c_int32),
    ]


def _sincos(angle):
    """b2Rot::Set(angle) as the reference evaluates it (libm sinf / cosf): the library's own host-side restatement."""
    from .batch import lib
    x, s, c = (np.array([angle], F32), np.zeros(1, F32), np.zeros(1, F32))
    lib().b2gpu_host_sincosf(x.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), 1)
    return s[0], c[0]


class Scene:
    """One world, as built by the user.  Mirrors b2World / b2Body construction calls."""

    def __init__(self, gravity=(0.0, -10.0), flags=WORLD_DEFAULT):
        self.gravity = (float(gravity[0]), float(gravity[1]))
        self.flags = int(flags)
        self.bodies: list[BodyDef] = []
        self.fixtures: list[FixtureDef] = []
        self.vertices: list[float] = []
        self.joints: list[JointDef] = []
        self._c = None

    # ---- b2World::CreateJoint (Box2D/Dynamics/b2World.cpp:211-270) ----
    def _local(self, body, point):
        """b2Body::GetLocalPoint of a freshly created body (b2MulT(xf, p)) in float32."""
        bd = self.bodies[body]
        s, c = _sincos(bd.angle)
        px, py = F32(F32(point[0]) - F32(bd.position[0])), F32(F32(point[1]) - F32(bd.position[1]))
        return float(F32(c * px) + F32(s * py)), float(F32(-s * px) + F32(c * py))

    def _joint(self, type, body_a, body_b, collide_connected):
        jd = JointDef()
        jd.type, jd.bodyA, jd.bodyB, jd.collideConnected = type, body_a, body_b, int(collide_connected)
        self.joints.append(jd)
        self._c = None
        return jd

    def create_revolute_joint(self, body_a, body_b, anchor, lower=0.0, upper=0.0, enable_limit=False,
                              max_motor_torque=0.0, motor_speed=0.0, enable_motor=False, collide_connected=False) -> int:
        """b2RevoluteJointDef::Initialize(bodyA, bodyB, anchor) + the def's members (b2RevoluteJoint.cpp:38-45)."""
        jd = self._joint(JOINT_REVOLUTE, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor)
        jd.referenceAngle = float(F32(self.bodies[body_b].angle) - F32(self.bodies[body_a].angle))
        jd.lower, jd.upper, jd.enableLimit = lower, upper, int(enable_limit)
        jd.maxMotor, jd.motorSpeed, jd.enableMotor = max_motor_torque, motor_speed, int(enable_motor)
        return len(self.joints) - 1

    def create_prismatic_joint(self, body_a, body_b, anchor, axis, lower=0.0, upper=0.0, enable_limit=False,
                               max_motor_force=0.0, motor_speed=0.0, enable_motor=False, collide_connected=False) -> int:
        """b2PrismaticJointDef::Initialize(bodyA, bodyB, anchor, axis) (b2PrismaticJoint.cpp:90-98)."""
        jd = self._joint(JOINT_PRISMATIC, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor)
        bd = self.bodies[body_a]  # GetLocalVector(axis) = b2MulT(q, axis)
        s, c = _sincos(bd.angle)
        ax, ay = F32(axis[0]), F32(axis[1])
        jd.localAxisA[0], jd.localAxisA[1] = float(F32(c * ax) + F32(s * ay)), float(F32(-s * ax) + F32(c * ay))
        jd.referenceAngle = float(F32(self.bodies[body_b].angle) - F32(self.bodies[body_a].angle))
        jd.lower, jd.upper, jd.enableLimit = lower, upper, int(enable_limit)
        jd.maxMotor, jd.motorSpeed, jd.enableMotor = max_motor_force, motor_speed, int(enable_motor)
        return len(self.joints) - 1

    def create_wheel_joint(self, body_a, body_b, anchor, axis, frequency_hz=2.0, damping_ratio=0.7,
                           max_motor_torque=0.0, motor_speed=0.0, enable_motor=False, collide_connected=False) -> int:
        """b2WheelJointDef::Initialize(bodyA, bodyB, anchor, axis) (b2WheelJoint.cpp:39-46)."""
        jd = self._joint(JOINT_WHEEL, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor)
        s, c = _sincos(self.bodies[body_a].angle)
        ax, ay = F32(axis[0]), F32(axis[1])
        jd.localAxisA[0], jd.localAxisA[1] = float(F32(c * ax) + F32(s * ay)), float(F32(-s * ax) + F32(c * ay))
        jd.frequencyHz, jd.dampingRatio = frequency_hz, damping_ratio
        jd.maxMotor, jd.motorSpeed, jd.enableMotor = max_motor_torque, motor_speed, int(enable_motor)
        return len(self.joints) - 1

    def create_weld_joint(self, body_a, body_b, anchor, frequency_hz=0.0, damping_ratio=0.0, collide_connected=False) -> int:
        """b2WeldJointDef::Initialize(bodyA, bodyB, anchor) (b2WeldJoint.cpp:37-44)."""
        jd = self._joint(JOINT_WELD, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor)
        jd.referenceAngle = float(F32(self.bodies[body_b].angle) - F32(self.bodies[body_a].angle))
        jd.frequencyHz, jd.dampingRatio = frequency_hz, damping_ratio
        return len(self.joints) - 1

    def create_rope_joint(self, body_a, body_b, anchor_a, anchor_b, max_length, collide_connected=False) -> int:
        """b2RopeJointDef (b2RopeJoint.h:28-52): local anchors from world points, maxLength."""
        jd = self._joint(JOINT_ROPE, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor_a)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor_b)
        jd.length = max_length
        return len(self.joints) - 1

    def create_friction_joint(self, body_a, body_b, anchor, max_force, max_torque, collide_connected=False) -> int:
        """b2FrictionJointDef::Initialize(bodyA, bodyB, anchor) (b2FrictionJoint.cpp:36-42) + maxForce / maxTorque."""
        jd = self._joint(JOINT_FRICTION, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor)
        jd.length, jd.maxMotor = max_force, max_torque
        return len(self.joints) - 1

    def create_motor_joint(self, body_a, body_b, linear_offset=None, angular_offset=None, max_force=1.0, max_torque=1.0,
                           correction_factor=0.3, collide_connected=False) -> int:
        """b2MotorJointDef::Initialize(bodyA, bodyB) (b2MotorJoint.cpp:36-45) unless offsets are given."""
        jd = self._joint(JOINT_MOTOR, body_a, body_b, collide_connected)
        if linear_offset is None:
            linear_offset = self._local(body_a, self.bodies[body_b].position)
        if angular_offset is None:
            angular_offset = float(F32(self.bodies[body_b].angle) - F32(self.bodies[body_a].angle))
        jd.localAnchorA[0], jd.localAnchorA[1] = linear_offset
        jd.referenceAngle = angular_offset
        jd.length, jd.maxMotor, jd.frequencyHz = max_force, max_torque, correction_factor
        return len(self.joints) - 1

    def create_mouse_joint(self, body_a, body_b, target, max_force, frequency_hz=5.0, damping_ratio=0.7, collide_connected=False) -> int:
        """b2MouseJointDef (b2MouseJoint.h:29-55): drags the point of body_b that is at `target` now towards `target`
        (a soft constraint with a force cap); body_a is unused by the solver (the Testbed passes the ground)."""
        jd = self._joint(JOINT_MOUSE, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = target
        jd.length, jd.frequencyHz, jd.dampingRatio = max_force, frequency_hz, damping_ratio
        return len(self.joints) - 1

    def create_gear_joint(self, joint1, joint2, ratio=1.0, collide_connected=False) -> int:
        """b2GearJointDef (b2GearJoint.h:27-47): joint1 / joint2 are indices of revolute or prismatic joints created
        earlier; coordinate1 + ratio * coordinate2 stays constant.  bodyA / bodyB = the joints' second bodies
        (b2GearJoint.cpp:59-60, 92-93)."""
        a, b = self.joints[joint1].bodyB, self.joints[joint2].bodyB
        jd = self._joint(JOINT_GEAR, a, b, collide_connected)
        jd.enableLimit, jd.enableMotor, jd.motorSpeed = joint1, joint2, ratio
        return len(self.joints) - 1

    def create_pulley_joint(self, body_a, body_b, ground_a, ground_b, anchor_a, anchor_b, ratio=1.0, collide_connected=True) -> int:
        """b2PulleyJointDef::Initialize(bodyA, bodyB, groundA, groundB, anchorA, anchorB, ratio) (b2PulleyJoint.cpp:37-56);
        collideConnected defaults to true like b2PulleyJointDef's constructor."""
        jd = self._joint(JOINT_PULLEY, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor_a)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor_b)
        jd.localAxisA[0], jd.localAxisA[1] = ground_a
        jd.lower, jd.upper = ground_b

        def dist(p, q):
            dx, dy = F32(F32(p[0]) - F32(q[0])), F32(F32(p[1]) - F32(q[1]))
            return float(np.sqrt(F32(dx * dx) + F32(dy * dy), dtype=F32))

        jd.length, jd.maxMotor, jd.motorSpeed = dist(anchor_a, ground_a), dist(anchor_b, ground_b), ratio
        return len(self.joints) - 1

    def create_distance_joint(self, body_a, body_b, anchor_a, anchor_b, frequency_hz=0.0, damping_ratio=0.0,
                              length=None, collide_connected=False) -> int:
        """b2DistanceJointDef::Initialize(bodyA, bodyB, anchorA, anchorB) (b2DistanceJoint.cpp:40-49)."""
        jd = self._joint(JOINT_DISTANCE, body_a, body_b, collide_connected)
        jd.localAnchorA[0], jd.localAnchorA[1] = self._local(body_a, anchor_a)
        jd.localAnchorB[0], jd.localAnchorB[1] = self._local(body_b, anchor_b)
        if length is None:
            dx, dy = F32(F32(anchor_b[0]) - F32(anchor_a[0])), F32(F32(anchor_b[1]) - F32(anchor_a[1]))
            length = float(np.sqrt(F32(dx * dx) + F32(dy * dy), dtype=F32))
        jd.length, jd.frequencyHz, jd.dampingRatio = length, frequency_hz, damping_ratio
        return len(self.joints) - 1

    # ---- b2World::CreateBody ----
    def create_body(self, type=STATIC_BODY, position=(0.0, 0.0), angle=0.0, linear_velocity=(0.0, 0.0),
                    angular_velocity=0.0, linear_damping=0.0, angular_damping=0.0, gravity_scale=1.0,
                    flags=BD_DEFAULT) -> int:
        bd = BodyDef()
        bd.type = type
        bd.flags = flags
        bd.position[0], bd.position[1] = position
        bd.angle = angle
        bd.linearVelocity[0], bd.linearVelocity[1] = linear_velocity
        bd.angularVelocity = angular_velocity
        bd.linearDamping = linear_damping
        bd.angularDamping = angular_damping
        bd.gravityScale = gravity_scale
        self.bodies.append(bd)
        self._c = None
        return len(self.bodies) - 1

    def _fixture(self, body, shape_type, kind, density, friction, restitution, sensor, category, mask, group):
        fd = FixtureDef()
        fd.body = body
        fd.shapeType = shape_type
        fd.shapeKind = kind
        fd.isSensor = int(sensor)
        fd.friction = friction
        fd.restitution = restitution
        fd.density = density
        fd.categoryBits = category
        fd.maskBits = mask
        fd.groupIndex = group
        self.fixtures.append(fd)
        self._c = None
        return fd

    _kw = dict(density=0.0, friction=0.2, restitution=0.0, sensor=False, category=1, mask=0xFFFF, group=0)

    # ---- b2Body::CreateFixture with each shape class ----
    def create_box(self, body, hx, hy, center=None, angle=0.0, **kw):
        k = {**self._kw, **kw}
        fd = self._fixture(body, SHAPE_POLYGON, POLY_BOX if center is None else POLY_BOX_OFFSET, k["density"],
                           k["friction"], k["restitution"], k["sensor"], k["category"], k["mask"], k["group"])
        fd.p[0], fd.p[1] = hx, hy
        if center is not None:
            fd.center[0], fd.center[1] = center
            fd.angle = angle
        return len(self.fixtures) - 1

    def create_polygon(self, body, points, **kw):
        k = {**self._kw, **kw}
        fd = self._fixture(body, SHAPE_POLYGON, POLY_SET, k["density"], k["friction"], k["restitution"],
                           k["sensor"], k["category"], k["mask"], k["group"])
        fd.firstVertex = len(self.vertices) // 2
        fd.vertexCount = len(points)
        for x, y in points:
            self.vertices += [float(x), float(y)]
        return len(self.fixtures) - 1

    def create_circle(self, body, radius, p=(0.0, 0.0), **kw):
        k = {**self._kw, **kw}
        fd = self._fixture(body, SHAPE_CIRCLE, 0, k["density"], k["friction"], k["restitution"], k["sensor"],
                           k["category"], k["mask"], k["group"])
        fd.radius = radius
        fd.p[0], fd.p[1] = p
        return len(self.fixtures) - 1

    def create_edge(self, body, v1, v2, v0=None, v3=None, **kw):
        k = {**self._kw, **kw}
        fd = self._fixture(body, SHAPE_EDGE, 0, k["density"], k["friction"], k["restitution"], k["sensor"],
                           k["category"], k["mask"], k["group"])
        fd.firstVertex = len(self.vertices) // 2
        fd.vertexCount = 2
        self.vertices += [float(v1[0]), float(v1[1]), float(v2[0]), float(v2[1])]
        if v0 is not None:
            fd.hasVertex0 = 1
            fd.vertex0[0], fd.vertex0[1] = v0
        if v3 is not None:
            fd.hasVertex3 = 1
            fd.vertex3[0], fd.vertex3[1] = v3
        return len(self.fixtures) - 1

    def create_chain(self, body, points, loop=False, prev=None, next=None, **kw):
        k = {**self._kw, **kw}
        fd = self._fixture(body, SHAPE_CHAIN, CHAIN_LOOP if loop else CHAIN_OPEN, k["density"], k["friction"],
                           k["restitution"], k["sensor"], k["category"], k["mask"], k["group"])
        fd.firstVertex = len(self.vertices) // 2
        fd.vertexCount = len(points)
        for x, y in points:
            self.vertices += [float(x), float(y)]
        if prev is not None:
            fd.hasVertex0 = 1
            fd.vertex0[0], fd.vertex0[1] = prev
        if next is not None:
            fd.hasVertex3 = 1
            fd.vertex3[0], fd.vertex3[1] = next
        return len(self.fixtures) - 1

    # ---- export ----
    @property
    def body_count(self):
        return len(self.bodies)

    @property
    def dynamic_body_count(self):
        return sum(1 for b in self.bodies if b.type == DYNAMIC_BODY)

    def c_scene(self) -> CScene:
        if self._c is None:
            nb, nf, nv = len(self.bodies), len(self.fixtures), len(self.vertices) // 2
            bodies = (BodyDef * max(nb, 1))(*self.bodies)
            fixtures = (FixtureDef * max(nf, 1))(*self.fixtures)
            verts = (C.c_float * max(2 * nv, 2))(*self.vertices)
            cs = CScene()
            cs.gravity[0], cs.gravity[1] = self.gravity
            cs.worldFlags = self.flags
            cs.bodyCount, cs.fixtureCount, cs.vertexCount = nb, nf, nv
            cs.bodies = C.cast(bodies, C.POINTER(BodyDef))
            cs.fixtures = C.cast(fixtures, C.POINTER(FixtureDef))
            cs.vertices = C.cast(verts, C.POINTER(C.c_float))
            nj = len(self.joints)
            joints = (JointDef * max(nj, 1))(*self.joints)
            cs.joints = C.cast(joints, C.POINTER(JointDef))
            cs.jointCount = nj
            self._c = (cs, bodies, fixtures, verts, joints)  # keep the arrays alive
        return self._c[0]


def scene_array(scenes):
    """Pack scenes into a contiguous C array of b2g_scene (keeps the Python objects alive)."""
    arr = (CScene * len(scenes))()
    for i, s in enumerate(scenes):
        cs = s.c_scene()
        C.memmove(C.byref(arr, i * C.sizeof(CScene)), C.byref(cs), C.sizeof(CScene))
    arr._keep = list(scenes)
    return arr


# --------------------------------------------------------------------------------------------
# Deterministic 32-bit LCG used by every randomised scene (documented in DESIGN.md):
#   state <- state * 1664525 + 1013904223 (mod 2^32);  u = (state >> 8) / 2^24 in [0, 1)
# --------------------------------------------------------------------------------------------
class LCG:
    def __init__(self, seed):
        self.state = (int(seed) * 2654435761 + 12345) & 0xFFFFFFFF

    def next_u(self) -> float:
        self.state = (self.state * 1664525 + 1013904223) & 0xFFFFFFFF
        return (self.state >> 8) / float(1 << 24)

    def uniform(self, lo, hi) -> float:
        return float(np.float32(lo + (hi - lo) * self.next_u()))


def hello_world(flags=WORLD_DEFAULT) -> Scene:
    """BASELINE config 1 — HelloWorld/HelloWorld.cpp:36-80."""
    s = Scene((0.0, -10.0), flags)
    ground = s.create_body(STATIC_BODY, (0.0, -10.0))
    s.create_box(ground, 50.0, 10.0)  # CreateFixture(&groundBox, 0.0f)
    body = s.create_body(DYNAMIC_BODY, (0.0, 4.0))
    s.create_box(body, 1.0, 1.0, density=1.0, friction=0.3)
    return s


def pyramid(rows=20, flags=WORLD_DEFAULT, ground="box", jitter_seed=None, jitter=0.02) -> Scene:
    """BASELINE configs 2/3/5 — Box2D Testbed "Pyramid" layout restated in SURVEY.md §8(d) C2.

    ground="box": variant A, SetAsBox(50, 10, center (0,-10), 0) on a static body at the origin.
    ground="edge": variant B, b2EdgeShape (-40,0)-(40,0).
    jitter_seed: world index w; every box's x gets U(-jitter, jitter) from LCG(w) (C3 "jittered").
    """
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    if ground == "box":
        s.create_box(g, 50.0, 10.0, center=(0.0, -10.0), angle=0.0)
    else:
        s.create_edge(g, (-40.0, 0.0), (40.0, 0.0))
    rng = LCG(jitter_seed) if jitter_seed is not None else None
    f32 = np.float32
    a = f32(0.5)
    x = np.array([-7.0, 0.75], dtype=f32)
    dx = np.array([0.5625, 1.25], dtype=f32)
    dy = np.array([1.125, 0.0], dtype=f32)
    for i in range(rows):
        y = x.copy()
        for _ in range(i, rows):
            px = float(y[0])
            if rng is not None:
                px = float(f32(px) + f32(rng.uniform(-jitter, jitter)))
            b = s.create_body(DYNAMIC_BODY, (px, float(y[1])))
            s.create_box(b, float(a), float(a), density=5.0)
            y = y + dy
        x = x + dx
    return s


def mixed(seed: int, flags=WORLD_DEFAULT, bodies=40) -> Scene:
    """BASELINE config 4 — SURVEY.md §8(d) C4: chain terrain + alternating circles / random convex polygons.

    Static body with b2ChainShape::CreateChain of 33 vertices at x = -16..16, heights U(0,2) (the two end
    vertices are raised to y = 8 so nothing rolls off); `bodies` dynamic bodies in columns 3 m apart
    (+U(0,1)) dropped from y in [6, 15]: even k circles r in [0.3,0.8], odd k convex polygons with 3-8 points
    on a jittered circle r in [0.4,0.9] through b2PolygonShape::Set; density 1.  All randomness from LCG(seed).
    """
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    pts = [(float(-16 + i), rng.uniform(0.0, 2.0)) for i in range(33)]
    pts[0] = (pts[0][0], 8.0)
    pts[32] = (pts[32][0], 8.0)
    s.create_chain(g, pts)
    for k in range(bodies):
        px = float(F32(-13.5 + 3.0 * (k % 10)) + F32(rng.uniform(0.0, 1.0)))
        py = float(F32(6.0 + 2.5 * (k // 10)) + F32(rng.uniform(0.0, 1.5)))
        b = s.create_body(DYNAMIC_BODY, (px, py))
        if k % 2 == 0:
            s.create_circle(b, rng.uniform(0.3, 0.8), density=1.0)
        else:
            n = 3 + int(rng.next_u() * 6) % 6
            r = rng.uniform(0.4, 0.9)
            poly = []
            for j in range(n):
                ang = (j + 0.8 * rng.next_u()) * (2.0 * np.pi / n)
                poly.append((float(F32(r * np.cos(ang))), float(F32(r * np.sin(ang)))))
            s.create_polygon(b, poly, density=1.0)
    return s


def bullets(seed: int, flags=WORLD_DEFAULT, shots=6) -> Scene:
    """Continuous-collision scene: bullet bodies fired at a thin static wall and at a stack of
    boxes, plus a kinematic paddle.  Exercises every branch of b2World::SolveTOI — bullet-vs-dynamic times of
    impact, sweeps at different alpha0, mini-islands with kinematic / static / bullet neighbours.
    All randomness from LCG(seed)."""
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    s.create_box(g, 40.0, 1.0, center=(0.0, -1.0), angle=0.0)
    s.create_box(g, 0.05, 4.0, center=(12.0, 4.0), angle=0.0)          # thin wall
    s.create_edge(g, (-20.0, 0.0), (-20.0, 12.0))                      # edge wall on the left
    for k in range(6):                                                 # stack
        b = s.create_body(DYNAMIC_BODY, (4.0 + rng.uniform(-0.05, 0.05), 0.5 + 1.0 * k))
        s.create_box(b, 0.5, 0.5, density=2.0, friction=0.4)
    paddle = s.create_body(KINEMATIC_BODY, (-6.0, 3.0), linear_velocity=(2.5, 0.0), angular_velocity=1.0)
    s.create_box(paddle, 1.5, 0.2)
    for k in range(shots):
        y = 0.6 + rng.uniform(0.0, 5.0)
        speed = 60.0 + rng.uniform(0.0, 120.0)
        left = k % 3 == 2
        b = s.create_body(DYNAMIC_BODY, (-15.0 + 2.0 * k if not left else 9.0, y),
                          linear_velocity=(-speed if left else speed, rng.uniform(-5.0, 5.0)),
                          angular_velocity=rng.uniform(-20.0, 20.0), flags=BD_DEFAULT | BD_BULLET)
        if k % 2 == 0:
            s.create_circle(b, rng.uniform(0.1, 0.25), density=8.0, restitution=0.3)
        else:
            s.create_box(b, rng.uniform(0.1, 0.3), rng.uniform(0.05, 0.2), density=8.0, friction=0.3)
    return s


def sensors(seed: int, flags=WORLD_DEFAULT, bodies=14) -> Scene:
    """Sensor scene (the sensor branch of b2Contact::Update): static sensor zones (box, circle,
    and one on an edge-ground body) that dynamic circles / boxes fall through, and dynamic bodies that carry a
    sensor halo next to their solid fixture, so sensor-vs-solid and sensor-vs-sensor pairs both occur.
    All randomness from LCG(seed)."""
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    s.create_edge(g, (-20.0, 0.0), (20.0, 0.0))
    s.create_box(g, 5.0, 1.0, center=(-4.0, 4.0), angle=0.2, sensor=True)       # tilted sensor slab
    s.create_circle(g, 2.5, p=(6.0, 3.0), sensor=True)                          # sensor disc
    s.create_edge(g, (-12.0, 1.5), (12.0, 1.5), sensor=True)                    # sensor trip wire
    z = s.create_body(KINEMATIC_BODY, (0.0, 7.0), linear_velocity=(1.0, 0.0), angular_velocity=0.5)
    s.create_box(z, 3.0, 0.5, sensor=True)                                      # moving sensor
    for k in range(bodies):
        px = float(F32(-9.0 + 1.5 * k) + F32(rng.uniform(0.0, 0.5)))
        py = float(F32(6.0 + 1.2 * (k % 4)) + F32(rng.uniform(0.0, 1.0)))
        b = s.create_body(DYNAMIC_BODY, (px, py), angle=rng.uniform(-1.0, 1.0))
        if k % 3 == 0:
            s.create_circle(b, rng.uniform(0.25, 0.6), density=1.0, restitution=0.2)
        else:
            s.create_box(b, rng.uniform(0.25, 0.6), rng.uniform(0.25, 0.6), density=1.0, friction=0.3)
        if k % 2 == 0:
            s.create_circle(b, rng.uniform(0.9, 1.6), sensor=True)              # halo around the solid fixture
    return s


def filtered(seed: int, flags=WORLD_DEFAULT, bodies=36) -> Scene:
    """Collision filtering and per-body dynamics parameters that every other scene leaves at their defaults
    (SURVEY.md §8 a21, a13): b2Filter categoryBits / maskBits / groupIndex — positive groups (always collide), negative
    groups (never collide) and one-sided masks, the rule of b2ContactFilter::ShouldCollide (b2WorldCallbacks.cpp:26-39)
    — and linearDamping / angularDamping (b2Island.cpp:217-218), gravityScale outside {0, 1}, fixedRotation bodies in
    contact.  Boxes, circles and polygons dropped in a heap onto a box ground with two static ledges that only some
    categories can stand on.  All randomness from LCG(seed)."""
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    s.create_box(g, 30.0, 1.0, center=(0.0, -1.0), angle=0.0, category=0x0010)                 # everyone lands here
    s.create_box(g, 4.0, 0.25, center=(-6.0, 3.0), angle=0.0, category=0x0002, mask=0x0004)    # ledge for category 4 only
    s.create_box(g, 4.0, 0.25, center=(6.0, 3.5), angle=0.1, category=0x0008, mask=0xFFFF, group=-3)  # not for group -3
    cats = [0x0001, 0x0002, 0x0004, 0x0008]
    for k in range(bodies):
        px = float(F32(-11.0 + 2.0 * (k % 12)) + F32(rng.uniform(0.0, 0.8)))
        py = float(F32(5.0 + 1.6 * (k // 12)) + F32(rng.uniform(0.0, 0.8)))
        bflags = BD_DEFAULT | (BD_FIXED_ROTATION if k % 5 == 0 else 0)
        b = s.create_body(DYNAMIC_BODY, (px, py), angle=rng.uniform(-0.6, 0.6), angular_velocity=rng.uniform(-2.0, 2.0),
                          linear_damping=(0.0, 0.3, 1.5)[k % 3], angular_damping=(0.0, 0.8)[k % 2],
                          gravity_scale=(1.0, 0.35, 2.5, -0.15, 1.0)[k % 5], flags=bflags)
        cat = cats[int(rng.next_u() * 4) % 4]
        u = rng.next_u()
        mask = 0xFFFF if u < 0.4 else (0xFFFF & ~cats[int(rng.next_u() * 4) % 4] if u < 0.8 else (cat | 0x0010))
        v = rng.next_u()
        group = 0 if v < 0.5 else (2 if v < 0.65 else (-3 if v < 0.85 else -7))
        kw = dict(density=1.0, friction=0.4, restitution=(0.0, 0.3)[k % 2], category=cat, mask=mask, group=group)
        if k % 3 == 0:
            s.create_circle(b, rng.uniform(0.3, 0.55), **kw)
        elif k % 3 == 1:
            s.create_box(b, rng.uniform(0.3, 0.6), rng.uniform(0.25, 0.5), **kw)
        else:
            n = 3 + int(rng.next_u() * 4) % 4
            r = rng.uniform(0.35, 0.6)
            poly = []
            for j in range(n):
                ang = (j + 0.7 * rng.next_u()) * (2.0 * np.pi / n)
                poly.append((float(F32(r * np.cos(ang))), float(F32(r * np.sin(ang)))))
            s.create_polygon(b, poly, **kw)
    return s


def jointed(seed: int, flags=WORLD_DEFAULT) -> Scene:
    """Joint scene: every mechanism the device path solves, in one world so that islands mix joints
    and contacts —
      * a plank bridge (revolute joints, no limits) between two posts with boxes dropped on it (Testbed "Bridge"),
      * a cart on a rail (prismatic joint with limits and a motor) carrying a pole (revolute joint): cart-pole,
      * a two-legged walker (revolute joints with limits and motors, collideConnected off) standing on the ground,
      * pendulums on rigid and soft distance joints, one pair of bodies tied by a distance joint with collideConnected.
    All randomness from LCG(seed)."""
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    s.create_edge(g, (-40.0, 0.0), (40.0, 0.0))
    # ---- bridge ----
    n = 10
    prev = g
    x0, y0 = -18.0, 6.0
    for i in range(n):
        b = s.create_body(DYNAMIC_BODY, (x0 + 0.5 + 1.0 * i, y0))
        s.create_box(b, 0.5, 0.125, density=20.0, friction=0.2)
        s.create_revolute_joint(prev, b, (x0 + 1.0 * i, y0))
        prev = b
    s.create_revolute_joint(prev, g, (x0 + 1.0 * n, y0))
    for k in range(3):
        b = s.create_body(DYNAMIC_BODY, (x0 + 3.0 + 2.0 * k + rng.uniform(-0.2, 0.2), 8.0 + 1.5 * k), angle=rng.uniform(-0.5, 0.5))
        if k % 2:
            s.create_circle(b, 0.4, density=1.0)
        else:
            s.create_box(b, 0.4, 0.4, density=1.0)
    # ---- cart-pole ----
    cart = s.create_body(DYNAMIC_BODY, (0.0, 3.0))
    s.create_box(cart, 0.6, 0.25, density=2.0)
    s.create_prismatic_joint(g, cart, (0.0, 3.0), (1.0, 0.0), lower=-3.0, upper=3.0, enable_limit=True,
                             max_motor_force=60.0, motor_speed=rng.uniform(1.0, 4.0) * (1 if seed % 2 else -1), enable_motor=True)
    pole = s.create_body(DYNAMIC_BODY, (0.0, 4.0), angle=rng.uniform(-0.3, 0.3))
    s.create_box(pole, 0.08, 1.0, density=1.0)
    s.create_revolute_joint(cart, pole, (0.0, 3.0))
    # ---- walker: torso + two legs of two segments ----
    tx, ty = 10.0 + rng.uniform(-0.5, 0.5), 2.6
    torso = s.create_body(DYNAMIC_BODY, (tx, ty))
    s.create_box(torso, 0.7, 0.3, density=3.0)
    for side in (-1, 1):
        hipx = tx + 0.4 * side
        thigh = s.create_body(DYNAMIC_BODY, (hipx, ty - 0.7))
        s.create_box(thigh, 0.12, 0.5, density=1.0, friction=0.8)
        s.create_revolute_joint(torso, thigh, (hipx, ty - 0.25), lower=-0.6, upper=0.8, enable_limit=True,
                                max_motor_torque=40.0, motor_speed=1.5 * side, enable_motor=True)
        shin = s.create_body(DYNAMIC_BODY, (hipx, ty - 1.65))
        s.create_box(shin, 0.1, 0.5, density=1.0, friction=0.8)
        s.create_revolute_joint(thigh, shin, (hipx, ty - 1.2), lower=-1.2, upper=0.05, enable_limit=True,
                                max_motor_torque=20.0, motor_speed=-1.0 * side, enable_motor=(side > 0))
    # ---- pendulums on distance joints ----
    for k in range(4):
        ax = 18.0 + 2.5 * k
        bx = ax + 2.0 + rng.uniform(0.0, 0.5)
        bob = s.create_body(DYNAMIC_BODY, (bx, 9.0))
        s.create_circle(bob, 0.35, density=2.0)
        s.create_distance_joint(g, bob, (ax, 10.0), (float(F32(bx)), 9.0),
                                frequency_hz=(0.0 if k % 2 == 0 else 2.0 + k), damping_ratio=(0.0 if k % 2 == 0 else 0.3))
    a = s.create_body(DYNAMIC_BODY, (30.0, 1.0))
    s.create_box(a, 0.5, 0.5, density=1.0)
    b = s.create_body(DYNAMIC_BODY, (31.5, 1.0), linear_velocity=(-3.0, 0.0))
    s.create_box(b, 0.5, 0.5, density=1.0)
    s.create_distance_joint(a, b, (30.0, 1.0), (31.5, 1.0), collide_connected=True)
    return s


def vehicles(seed: int, flags=WORLD_DEFAULT) -> Scene:
    """Wheel and weld joints: a car (chassis + two wheels on b2WheelJoint springs, the rear one
    motorised) driving over bumpy chain terrain towards a cantilever of planks held by weld joints (rigid and soft),
    plus a welded compound that tumbles down, a weight on a rope joint and a puck slowed by a friction joint.
    All randomness from LCG(seed)."""
    rng = LCG(seed)
    s = Scene((0.0, -10.0), flags)
    g = s.create_body(STATIC_BODY, (0.0, 0.0))
    pts = [(-30.0, 0.0), (-10.0, 0.0)] + [(-10.0 + 2.0 * (i + 1), rng.uniform(0.0, 0.6)) for i in range(12)] + [(16.0, 0.0), (40.0, 0.0)]
    s.create_chain(g, pts)
    # ---- car (Testbed "Car" proportions) ----
    cx = -20.0 + rng.uniform(-1.0, 1.0)
    chassis = s.create_body(DYNAMIC_BODY, (cx, 1.0))
    s.create_polygon(chassis, [(-1.5, -0.5), (1.5, -0.5), (1.5, 0.0), (0.0, 0.9), (-1.15, 0.9), (-1.5, 0.2)], density=1.0)
    for i, (wx, motor) in enumerate(((-1.0, True), (1.0, False))):
        wheel = s.create_body(DYNAMIC_BODY, (cx + wx, 0.35))
        s.create_circle(wheel, 0.4, density=1.0, friction=0.9)
        s.create_wheel_joint(chassis, wheel, (cx + wx, 0.35), (0.0, 1.0), frequency_hz=(4.0 if i == 0 else rng.uniform(3.0, 5.0)),
                             damping_ratio=0.7, max_motor_torque=(20.0 if motor else 10.0), motor_speed=(-12.0 if motor else 0.0),
                             enable_motor=motor)
    # ---- cantilever: planks welded end to end, first to the ground ----
    prev = g
    for i in range(6):
        b = s.create_body(DYNAMIC_BODY, (20.0 + 0.5 + 1.0 * i, 5.0))
        s.create_box(b, 0.5, 0.125, density=20.0)
        s.create_weld_joint(prev, b, (20.0 + 1.0 * i, 5.0), frequency_hz=(0.0 if i < 3 else 5.0), damping_ratio=(0.0 if i < 3 else 0.7))
        prev = b
    # ---- welded compound dropped onto the terrain ----
    a = s.create_body(DYNAMIC_BODY, (2.0, 6.0), angle=rng.uniform(-0.8, 0.8), angular_velocity=rng.uniform(-3.0, 3.0))
    s.create_box(a, 0.6, 0.2, density=2.0)
    b = s.create_body(DYNAMIC_BODY, (2.9, 6.0))
    s.create_circle(b, 0.35, density=1.0)
    s.create_weld_joint(a, b, (2.6, 6.0))
    # ---- a weight on a rope (slack at first, then taut) and a puck held by a friction joint against the ground ----
    wgt = s.create_body(DYNAMIC_BODY, (-26.0 + rng.uniform(0.0, 0.5), 9.0))
    s.create_box(wgt, 0.3, 0.3, density=4.0)
    s.create_rope_joint(g, wgt, (-27.0, 12.0), (float(F32(s.bodies[wgt].position[0])), 9.0), max_length=4.5)
    puck = s.create_body(DYNAMIC_BODY, (-24.0, 3.0), linear_velocity=(rng.uniform(3.0, 6.0), 2.0), angular_velocity=5.0, gravity_scale=0.0)
    s.create_circle(puck, 0.3, density=1.0)
    s.create_friction_joint(g, puck, (-24.0, 3.0), max_force=rng.uniform(1.0, 3.0), max_torque=0.4)
    # ---- a platform driven towards an offset from the ground by a motor joint ----
    plat = s.create_body(DYNAMIC_BODY, (-32.0, 6.0), gravity_scale=0.0)
    s.create_box(plat, 1.0, 0.2, density=2.0)
    s.create_motor_joint(g, plat, linear_offset=(-30.0 + rng.uniform(-1.0, 1.0), 8.0), angular_offset=0.5, max_force=200.0,
                         max_torque=100.0, correction_factor=0.3)
    # ---- two weights over a pulley (Testbed "Pulleys"), ratio 1.5 ----
    pa = s.create_body(DYNAMIC_BODY, (-36.0, 5.0))
    s.create_box(pa, 0.5, 0.8, density=5.0)
    pb = s.create_body(DYNAMIC_BODY, (-34.0, 5.0 + rng.uniform(0.0, 0.5)))
    s.create_box(pb, 0.5, 0.8, density=5.0)
    s.create_pulley_joint(pa, pb, (-36.0, 11.0), (-34.0, 11.0), (-36.0, 5.8), (-34.0, float(F32(s.bodies[pb].position[1]) + F32(0.8))), ratio=1.5)
    # ---- gears (Testbed "Gears"): two wheels on ground pivots and a plank on a vertical slide, coupled by two gear joints ----
    r1, r2 = 0.5, 1.0
    w1 = s.create_body(DYNAMIC_BODY, (-44.0, 12.0), angular_velocity=rng.uniform(1.0, 3.0))
    s.create_circle(w1, r1, density=5.0)
    jw1 = s.create_revolute_joint(g, w1, (-44.0, 12.0))
    w2 = s.create_body(DYNAMIC_BODY, (-42.5, 12.0))
    s.create_circle(w2, r2, density=5.0)
    jw2 = s.create_revolute_joint(g, w2, (-42.5, 12.0))
    plank = s.create_body(DYNAMIC_BODY, (-41.0, 12.0 + rng.uniform(-0.2, 0.2)))
    s.create_box(plank, 0.5, 2.5, density=5.0)
    jpl = s.create_prismatic_joint(g, plank, (-41.0, 12.0), (0.0, 1.0), lower=-5.0, upper=5.0, enable_limit=True)
    s.create_gear_joint(jw1, jw2, ratio=r2 / r1)
    s.create_gear_joint(jw2, jpl, ratio=-1.0 / r2)
    # ---- a crate dragged by a mouse joint towards the point where it was grabbed ----
    crate = s.create_body(DYNAMIC_BODY, (-48.0, 9.0), linear_velocity=(rng.uniform(2.0, 5.0), 3.0), angular_velocity=2.0)
    s.create_box(crate, 0.5, 0.5, density=1.0)
    s.create_mouse_joint(g, crate, (-47.7, 9.3), max_force=1000.0 * 1.0)
    c2 = s.create_body(DYNAMIC_BODY, (2.0, 8.0), flags=BD_DEFAULT | BD_FIXED_ROTATION)
    s.create_box(c2, 0.3, 0.3, density=1.0)
    d2 = s.create_body(DYNAMIC_BODY, (2.6, 8.0), flags=BD_DEFAULT | BD_FIXED_ROTATION)
    s.create_box(d2, 0.3, 0.3, density=1.0)
    s.create_weld_joint(c2, d2, (2.3, 8.0))  # both bodies with fixed rotation: K.ez.z == 0 branch
    return s


class SceneBatch:
    """Many worlds that share fixtures/vertices and differ only in their body definitions.

    Built with numpy so that 10^5 jittered worlds do not cost 10^5 x Python object construction.
    ``c_array`` is a contiguous array of b2g_scene whose ``bodies`` pointers point into one big buffer.
    """

    def __init__(self, base: Scene, bodies: np.ndarray):
        self.base = base
        self.count = bodies.shape[0]
        self._bodies = np.ascontiguousarray(bodies)
        cs = base.c_scene()
        arr = (CScene * self.count)()
        stride = self._bodies.strides[0]
        addr = self._bodies.ctypes.data
        for i in range(self.count):
            C.memmove(C.byref(arr, i * C.sizeof(CScene)), C.byref(cs), C.sizeof(CScene))
            arr[i].bodies = C.cast(addr + i * stride, C.POINTER(BodyDef))
        self.c_array = arr

    def scene(self, i) -> Scene:
        """Materialise world i as a standalone Scene (for the oracles)."""
        s = Scene(self.base.gravity, self.base.flags)
        s.fixtures = self.base.fixtures
        s.vertices = self.base.vertices
        s.joints = self.base.joints
        raw = self._bodies[i].tobytes()
        n = len(self.base.bodies)
        s.bodies = list((BodyDef * n).from_buffer_copy(raw))
        return s


BODY_DTYPE = np.dtype([("type", "<i4"), ("flags", "<i4"), ("position", "<f4", 2), ("angle", "<f4"),
                       ("linearVelocity", "<f4", 2), ("angularVelocity", "<f4"), ("linearDamping", "<f4"),
                       ("angularDamping", "<f4"), ("gravityScale", "<f4")])
assert BODY_DTYPE.itemsize == C.sizeof(BodyDef)


def pyramid_batch(count, rows=20, flags=WORLD_DEFAULT, first_seed=0, jitter=0.02, jittered=True) -> SceneBatch:
    """`count` pyramid worlds; world i equals ``pyramid(rows, flags, jitter_seed=first_seed + i)`` bit for bit."""
    base = pyramid(rows, flags)
    nb = len(base.bodies)
    one = np.frombuffer(b"".join(bytes(b) for b in base.bodies), dtype=BODY_DTYPE)
    bodies = np.tile(one, (count, 1))
    if jittered:
        seeds = (np.arange(first_seed, first_seed + count, dtype=np.uint64) * 2654435761 + 12345) & 0xFFFFFFFF
        state = seeds.copy()
        for j in range(1, nb):  # body 0 is the ground; boxes in creation order
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
            u = (state >> 8).astype(np.float64) / float(1 << 24)
            delta = (-jitter + (2 * jitter) * u).astype(np.float32)
            bodies["position"][:, j, 0] = bodies["position"][:, j, 0] + delta
    return SceneBatch(base, bodies)
