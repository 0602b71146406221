/*
 * b2gpu.h — the drop-in boundary: a C ABI over the CUDA runtime for stepping a batch of
 * independent Box2D worlds on one B200.  Plain pointers and sizes only; no C++ or torch types.
 *
 * The reference (ilektron/box2dpp) has no FFI layer: its boundary is the C++ class b2World.  Each
 * entry point below names the reference interface it replaces; the C++ host class b2WorldBatch
 * (box2dpp_b200/include/Box2D/Gpu/b2WorldBatch.h) keeps the reference's spelling on top of this ABI,
 * and INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Two layers live in libb2gpu.so:
 *   b2gpu_batch_*   device layer: arenas, kernels, field upload/download, per-step I/O.
 *   b2gpu_scene_*   host layer: turns b2g_scene descriptions (the CreateBody/CreateFixture calls
 *                   of the reference, include/b2g_scene.h) into arena fields (mass data, hulls, fat
 *                   AABBs, proxy ids) exactly like b2Body/b2Fixture/b2PolygonShape do on the CPU.
 *
 * Error convention: every call returns B2GPU_OK (0) or a negative B2GPU_ERR_*; the message of the last
 * failure on the calling thread is b2gpu_last_error().  The caller owns every host buffer; device
 * memory is owned by the batch handle.  A handle is not thread-safe; distinct handles are.
 * There is NO CPU fallback: without a usable CUDA device b2gpu_batch_create fails with
 * B2GPU_ERR_NO_DEVICE.
 */
#ifndef B2GPU_H
#define B2GPU_H

#include <stddef.h>
#include <stdint.h>

#include "b2g_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    B2GPU_OK = 0,
    B2GPU_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product does not run on the CPU */
    B2GPU_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    B2GPU_ERR_INVALID = -3,     /* bad argument */
    B2GPU_ERR_UNSUPPORTED = -4, /* a value outside what the device path builds (e.g. a joint type that is none of b2JointType 1-11) */
    B2GPU_ERR_CAPACITY = -5     /* a per-world capacity in the descriptor is too small */
};

/* Event types in the readback buffer == b2ContactListener callbacks
 * (Box2D/Dynamics/b2WorldCallbacks.h:74-120; call sites b2Contact.cpp:242-255, b2ContactManager.cpp:49-52,
 * b2Island.cpp:517-540). */
enum { B2GPU_EVENT_BEGIN = 0, B2GPU_EVENT_END = 1, B2GPU_EVENT_PRESOLVE = 2, B2GPU_EVENT_POSTSOLVE = 3 };

/* Per-world status bits (b2gpu_batch_download_field(B2GPU_F_STATUS)). */
enum { B2GPU_STATUS_CONTACT_OVERFLOW = 1, B2GPU_STATUS_EVENT_OVERFLOW = 2, B2GPU_STATUS_BAD_COMMAND = 4,
       B2GPU_STATUS_SPAWN_OVERFLOW = 8 /* a body created after the first step did not fit the arenas */ };

/* Between-step body edits: one record per call the reference's user would make on a b2Body
 * between two Steps.  Box2D/Dynamics/b2Body.h:738-833 (Apply*), :408-467 (SetLinearVelocity, SetAngularVelocity,
 * SetAwake), Box2D/Dynamics/b2Body.cpp:418-440 (SetTransform). */
enum {
    B2GPU_CMD_APPLY_FORCE = 0,           /* v = force, p = world point        b2Body::ApplyForce(force, point, wake) */
    B2GPU_CMD_APPLY_FORCE_TO_CENTER = 1, /* v = force                         b2Body::ApplyForceToCenter(force, wake) */
    B2GPU_CMD_APPLY_TORQUE = 2,          /* v[0] = torque                     b2Body::ApplyTorque(torque, wake) */
    B2GPU_CMD_APPLY_LINEAR_IMPULSE = 3,  /* v = impulse, p = world point      b2Body::ApplyLinearImpulse(impulse, point, wake) */
    B2GPU_CMD_APPLY_ANGULAR_IMPULSE = 4, /* v[0] = impulse                    b2Body::ApplyAngularImpulse(impulse, wake) */
    B2GPU_CMD_SET_LINEAR_VELOCITY = 5,   /* v = velocity                      b2Body::SetLinearVelocity(v) */
    B2GPU_CMD_SET_ANGULAR_VELOCITY = 6,  /* v[0] = omega                      b2Body::SetAngularVelocity(omega) */
    B2GPU_CMD_SET_TRANSFORM = 7,         /* v = position, p[0] = angle        b2Body::SetTransform(position, angle) */
    B2GPU_CMD_SET_AWAKE = 8,             /* wake = flag                       b2Body::SetAwake(flag) */
    B2GPU_CMD_DESTROY_BODY = 9,          /*                                   b2World::DestroyBody(body), b2World.cpp:124-203:
                                            its joints, contacts (EndContact events, wake-ups) and proxies go with it;
                                            the body index stays reserved (later commands on it are BAD_COMMAND) */
    B2GPU_CMD_MOUSE_SET_TARGET = 10,     /* v = target                        b2MouseJoint::SetTarget(target) (b2MouseJoint.cpp:54-61) on the
                                            mouse joint whose bodyB is `body` (the newest one if it has several): wakes the
                                            body; BAD_COMMAND if the body carries no mouse joint */
    /* ---- edits that change what collides (all of them leave the rest to the next Step, like the reference) ---- */
    B2GPU_CMD_SET_FILTER = 11,           /* wake = fixture index, v[0] / v[1] = the BIT PATTERNS of two int32: categoryBits |
                                            maskBits << 16, groupIndex   b2Fixture::SetFilterData + Refilter (b2Fixture.cpp:185-215):
                                            the fixture's contacts are flagged for filtering (b2ContactManager::Collide re-asks
                                            ShouldCollide at the next Step and destroys those that fail, :124-146), its proxies are
                                            touched (b2BroadPhase::TouchProxy) so that new pairs can form */
    B2GPU_CMD_DESTROY_JOINT = 12,        /* wake = joint id (creation order in its world, as returned by b2gpu_batch_create_joints;
                                            scene joints are 0 .. jointCount-1), body ignored   b2World::DestroyJoint (b2World.cpp:267-360):
                                            wakes both bodies; if the joint had collideConnected == false the contacts between its
                                            two bodies are flagged for filtering */
    B2GPU_CMD_SET_TYPE = 13,             /* wake = b2BodyType          b2Body::SetType (b2Body.cpp:115-166): mass data reset, a body
                                            turned static is stopped and its proxies re-synchronised, the body wakes, forces are
                                            cleared, its contacts are destroyed (no EndContact there, like the reference: :150-157
                                            calls b2ContactManager::Destroy, which does report it), its proxies are touched */
    B2GPU_CMD_SET_ACTIVE = 14,           /* wake = flag                b2Body::SetActive (b2Body.cpp:455-510): proxies created /
                                            destroyed, contacts destroyed on deactivation */
    B2GPU_CMD_SET_MASS_DATA = 15,        /* v[0] = mass, v[1] = I (about the local origin), p = center   b2Body::SetMassData (:354-394) */
    B2GPU_CMD_RESET_MASS_DATA = 16,      /*                             b2Body::ResetMassData (:283-352) */
    B2GPU_CMD_DESTROY_FIXTURE = 17,      /* wake = fixture index       b2Body::DestroyFixture (:216-281): its contacts are destroyed
                                            (EndContact events), its proxies leave the broad phase, the mass data is reset */
    /* ---- plain setters (b2Body.h:590-690, b2Body.cpp:512-534, b2Fixture.h:300-345, b2Fixture.cpp:217-224) ---- */
    B2GPU_CMD_SET_DAMPING = 18,          /* v[0] = linear, v[1] = angular   b2Body::SetLinearDamping + SetAngularDamping */
    B2GPU_CMD_SET_GRAVITY_SCALE = 19,    /* v[0] = scale                b2Body::SetGravityScale */
    B2GPU_CMD_SET_BULLET = 20,           /* wake = flag                 b2Body::SetBullet */
    B2GPU_CMD_SET_SLEEPING_ALLOWED = 21, /* wake = flag                 b2Body::SetSleepingAllowed (false also wakes the body) */
    B2GPU_CMD_SET_FIXED_ROTATION = 22,   /* wake = flag                 b2Body::SetFixedRotation (angular velocity cleared, mass data reset) */
    B2GPU_CMD_SET_SENSOR = 23,           /* wake = fixture index, v[0] != 0 = sensor   b2Fixture::SetSensor (wakes the body; the fixture's
                                            contacts follow at their next Update, as the reference reads the flag there) */
    B2GPU_CMD_SET_MATERIAL = 24,         /* wake = fixture index, v[0] = friction, v[1] = restitution   b2Fixture::SetFriction +
                                            SetRestitution (existing contacts keep their mixed values, like the reference) */
    B2GPU_CMD_SET_DENSITY = 25,          /* wake = fixture index, v[0] = mass, v[1] = I, p = center: the fixture's b2MassData under
                                            its NEW density (b2Shape::ComputeMass, b2gpu_host_fixture_mass_data computes it)
                                            b2Fixture::SetDensity (b2Fixture.h:313-317): like the reference's call it does not
                                            touch the body's mass, B2GPU_CMD_RESET_MASS_DATA does.  All-zero mass data = density 0
                                            (b2Body::ResetMassData skips the fixture; for the massless edge / chain shapes the two
                                            are the same arithmetic) */
    B2GPU_CMD_COUNT
};
typedef struct b2gpu_body_command {
    int32_t world, body; /* world index in the batch, body index = creation order within the world */
    int32_t op;          /* B2GPU_CMD_* */
    int32_t wake;        /* the call's `wake` argument (SET_AWAKE: the flag) */
    float v[2];
    float p[2];
} b2gpu_body_command;

typedef struct b2gpu_batch b2gpu_batch;

/* Replaces the allocators a b2World embeds (b2BlockAllocator / b2StackAllocator members,
 * Box2D/Dynamics/b2World.h:254-256): fixed per-world capacities for the world-interleaved arenas. */
typedef struct b2gpu_batch_desc {
    int32_t device;        /* CUDA device ordinal */
    int32_t numWorlds;
    int32_t maxBodies;     /* per world */
    int32_t maxFixtures;
    int32_t maxProxies;
    int32_t maxContacts;   /* contact list capacity; overflow sets B2GPU_STATUS_CONTACT_OVERFLOW */
    int32_t maxShapeWords;
    int32_t maxChainVerts;
    int32_t maxEvents;     /* per world per step */
    int32_t eventMask;     /* bit (1 << B2GPU_EVENT_*) per recorded event type; 0 = no listener */
    int32_t blockSize;     /* threads per CTA, multiple of 32; 0 = default */
    int32_t reserved;      /* bit 0: force the generic solver kernel (body state in global memory);
                              bit 1: rebuild the islands every step (validation of the topology-dirty tracking);
                              bit 2: type-binned narrow phase (set by b2gpu_scene_capacities when a scene mixes shape
                                     types or has sensors; results do not depend on it) */
    int32_t maxJoints;     /* per world; > 0 selects the generic solver kernel (joints are solved there) */
} b2gpu_batch_desc;

/* Arena fields reachable through upload/download.  Host layout is world-major:
 * host[(world * capacity + item) * components + c].  Component type is float unless noted. */
enum {
    B2GPU_F_GRAVITY = 0,        /* [W][1][2]                 b2World::m_gravity */
    B2GPU_F_WORLD_FLAGS,        /* [W][1][1] int32           B2G_WORLD_* */
    B2GPU_F_INV_DT0,            /* [W][1][1]                 b2World::m_inv_dt0 */
    B2GPU_F_BODY_COUNT,         /* [W][1][1] int32 */
    B2GPU_F_PROXY_COUNT,        /* [W][1][1] int32 */
    B2GPU_F_CONTACT_COUNT,      /* [W][1][1] int32 */
    B2GPU_F_STATUS,             /* [W][1][1] int32           B2GPU_STATUS_* */
    B2GPU_F_EVENT_COUNT,        /* [W][1][1] int32 */
    B2GPU_F_ISLAND_COUNT,       /* [W][1][1] int32 */
    B2GPU_F_STEP_COUNT,         /* [W][1][1] int32 */
    B2GPU_F_BODY_XF,            /* [W][NB][4]  m_xf: p.x p.y q.s q.c */
    B2GPU_F_BODY_SWEEP,         /* [W][NB][4]  m_sweep.c.xy, m_sweep.a, m_sleepTime */
    B2GPU_F_BODY_SWEEP0,        /* [W][NB][4]  m_sweep.c0.xy, m_sweep.a0, 0 */
    B2GPU_F_BODY_VEL,           /* [W][NB][4]  m_linearVelocity.xy, m_angularVelocity, 0 */
    B2GPU_F_BODY_FORCE,         /* [W][NB][4]  m_force.xy, m_torque, 0 */
    B2GPU_F_BODY_MASS,          /* [W][NB][4]  m_invMass, m_invI, m_sweep.localCenter.xy */
    B2GPU_F_BODY_DAMP,          /* [W][NB][4]  m_linearDamping, m_angularDamping, m_gravityScale, m_mass */
    B2GPU_F_BODY_FLAGS,         /* [W][NB][1] int32: type | island 4 | awake 8 | autoSleep 16 | bullet 32 | fixedRotation 64 | active 128 */
    B2GPU_F_FIXTURE_INFO,       /* [W][NF][1] int32: body | shapeType << 16 | isSensor << 18 */
    B2GPU_F_FIXTURE_SHAPE,      /* [W][NF][1] int32: word offset into the shape table */
    B2GPU_F_FIXTURE_MAT,        /* [W][NF][2]  friction, restitution */
    B2GPU_F_FIXTURE_FILTER,     /* [W][NF][2] int32: categoryBits | maskBits << 16, groupIndex */
    B2GPU_F_PROXY_FAT,          /* [W][NP][4]  fat AABB lo.xy hi.xy */
    B2GPU_F_PROXY_INFO,         /* [W][NP][1] int32: fixture | child << 16 */
    B2GPU_F_PROXY_MOVED,        /* [W][NP][1] int32: 1 = in b2BroadPhase::m_moveBuffer */
    B2GPU_F_SHAPE_WORDS,        /* [W][NS][1]  shape table (layout: box2dpp_b200/csrc/b2g_layout.h) */
    B2GPU_F_CHAIN_VERTS,        /* [W][NV][2] */
    B2GPU_F_CONTACT_A,          /* [W][NC][4] int32: proxyA|proxyB<<16, bodyA|bodyB<<16, flags, type|pointCount<<8|shapeTypes<<16 */
    B2GPU_F_CONTACT_K,          /* [W][NC][4] int32: id.key of point 0, of point 1, friction bits, restitution bits */
    B2GPU_F_CONTACT_S,          /* [W][NC][2] int32 */
    B2GPU_F_CONTACT_P0,         /* [W][NC][4]  points[0]: localPoint.xy, normalImpulse, tangentImpulse */
    B2GPU_F_CONTACT_P1,         /* [W][NC][4]  points[1] */
    B2GPU_F_CONTACT_N,          /* [W][NC][4]  localNormal.xy, localPoint.xy */
    B2GPU_F_EVENT_A,            /* [W][NE][4] int32: type, fixtureA|childA<<16, fixtureB|childB<<16, count */
    B2GPU_F_EVENT_B,            /* [W][NE][4]  normalImpulses[2], tangentImpulses[2] */
    B2GPU_F_ISLAND_CONTACTS,    /* [W][NC][1] int32: contact array index in solver order (last step) */
    B2GPU_F_ISLAND_CONTACT_END, /* [W][NB][1] int32 */
    B2GPU_F_JOINT_COUNT,        /* [W][1][1] int32 */
    B2GPU_F_JOINT_I,            /* [W][NJ][4] int32: type | collideConnected << 8 | enableLimit << 9 | enableMotor << 10,
                                   bodyA | bodyB << 16, m_limitState, 0 */
    B2GPU_F_JOINT_F0,           /* [W][NJ][4]  localAnchorA.xy, localAnchorB.xy */
    B2GPU_F_JOINT_F1,           /* [W][NJ][4]  revolute / prismatic: referenceAngle, lower, upper, maxMotorTorque|Force;
                                               distance: length, frequencyHz, dampingRatio, 0;
                                               wheel: frequencyHz, dampingRatio, 0, maxMotorTorque;
                                               weld: referenceAngle, frequencyHz, dampingRatio, 0;
                                               rope: maxLength; friction: maxForce, maxTorque */
    B2GPU_F_JOINT_F2,           /* [W][NJ][4]  motorSpeed, localXAxisA.xy (normalised), 0 */
    B2GPU_F_JOINT_IMPULSE,      /* [W][NJ][4]  m_impulse.xyz, m_motorImpulse (warm-starting state);
                                               wheel: m_impulse, m_springImpulse, 0, m_motorImpulse */
    B2GPU_F_FIXTURE_COUNT,      /* [W][1][1] int32  fixture slots in use (destroyed ones included) */
    B2GPU_F_SHAPE_WORD_COUNT,   /* [W][1][1] int32  words of the shape table in use */
    B2GPU_F_CHAIN_VERT_COUNT,   /* [W][1][1] int32 */
    B2GPU_F_PROXY_FREE_COUNT,   /* [W][1][1] int32  freed proxy slots (b2DynamicTree's free list, LIFO) */
    B2GPU_F_PROXY_FREE,         /* [W][NP][1] int32 */
    B2GPU_F_NEW_FIXTURE,        /* [W][1][1] int32  b2World::e_newFixture: the next Step starts with FindNewContacts */
    B2GPU_F_JOINT_SEQ,          /* [W][1][1] int32  joints created so far = the next joint id (ids ride in JOINT_I.x bits 16-31) */
    B2GPU_F_FIXTURE_MASS,       /* [W][NF][4]  b2Fixture::GetMassData: mass, center.xy, I (what b2Body::ResetMassData sums) */
    B2GPU_F_COUNT
};

/* ---- device layer ------------------------------------------------------------------------------ */

/* Replaces b2World::b2World(gravity) for a whole batch (Box2D/Dynamics/b2World.cpp:39-55): allocates the
 * arenas on desc->device.  All counts start at zero. */
int b2gpu_batch_create(const b2gpu_batch_desc* desc, b2gpu_batch** out);

/* Replaces b2World::~b2World (b2World.cpp:57-76). */
int b2gpu_batch_destroy(b2gpu_batch* batch);

/* Copy a field for worlds [firstWorld, firstWorld + worldCount) between a world-major host buffer and the
 * interleaved arena.  Replaces direct member access / getters of b2Body, b2Fixture, b2Contact
 * (Box2D/Dynamics/b2Body.h:163-197, Contacts/b2Contact.h:92-139) and doubles as snapshot / restore. */
int b2gpu_batch_upload_field(b2gpu_batch* batch, int32_t field, const void* host, int32_t firstWorld, int32_t worldCount);
int b2gpu_batch_download_field(b2gpu_batch* batch, int32_t field, void* host, int32_t firstWorld, int32_t worldCount);
/* Capacity (items per world) and components per item of a field. */
int b2gpu_batch_field_shape(b2gpu_batch* batch, int32_t field, int32_t* capacity, int32_t* components);

/* Replaces b2World::Step(dt, velocityIterations, positionIterations) (b2World.cpp:888-954) for every world
 * of the batch.  Asynchronous on the batch's stream; b2gpu_batch_sync waits for it. */
int b2gpu_batch_step(b2gpu_batch* batch, float dt, int32_t velocityIterations, int32_t positionIterations);
/* Same, `steps` times, replayed from a captured CUDA graph. */
int b2gpu_batch_step_n(b2gpu_batch* batch, float dt, int32_t velocityIterations, int32_t positionIterations, int32_t steps);
int b2gpu_batch_sync(b2gpu_batch* batch);

/* Per-step I/O for a caller that drives the batch (host buffers, pinned or not):
 * apply_forces replaces b2Body::ApplyForceToCenter / ApplyTorque (Box2D/Dynamics/b2Body.h:738-790) for every
 * body: forces[(world * maxBodies + body) * 3] = fx, fy, torque; non-zero entries wake the body.
 * read_transforms replaces b2Body::GetPosition / GetAngle / GetTransform (b2Body.h:500-533):
 * out[(world * maxBodies + body) * 4] = p.x, p.y, q.s, q.c. */
int b2gpu_batch_apply_forces(b2gpu_batch* batch, const float* forces);
int b2gpu_batch_read_transforms(b2gpu_batch* batch, float* out);

/* One step with its per-step I/O fused in: apply_forces(forces) + b2World::Step(dt, velIters, posIters)
 * (Box2D/Dynamics/b2World.cpp:888-954) + read_transforms(transforms) for every world, the loop body of a caller that
 * reads observations and writes controls every step.  Same results as the three separate calls (bit-identical);
 * the difference is scheduling: the batch is cut into world groups, each group copies its own slice of `forces`
 * in and of `transforms` out on its own stream, so the PCIe transfers of one group run under the kernels of the
 * others.  While no dynamic body of the batch can be asleep (no world allows sleeping, none was built or set asleep)
 * a force cannot wake anybody, so a group's forces are copied on a side stream under its contact update and island
 * build and applied just before the velocities are integrated — the first thing that reads them.  Either pointer may
 * be null (that side is skipped); pinned host buffers are needed for the overlap.
 * Synchronises: on return `transforms` holds this step's result. */
int b2gpu_batch_step_io(b2gpu_batch* batch, const float* forces, float dt, int32_t velIters, int32_t posIters,
                        float* transforms);

/* Which bodies the dense per-step I/O above carries, and in which format — so that only what the caller controls and
 * observes crosses the bus (e.g. not the static ground, three floats per pose instead of four).
 *   forceBodies[forceCount]: body indices, the same in every world; `forces` of apply_forces / step_io then is
 *       forces[(world * forceCount + k) * 3] for body forceBodies[k].  forceCount -1: every body slot (the default), 0: none.
 *   poseBodies[poseCount]: likewise for `out` / `transforms`: out[(world * poseCount + k) * F].
 *   poseFormat: B2GPU_POSE_XF   p.x p.y q.s q.c (b2Body::GetTransform, F = 4, the default)
 *               B2GPU_POSE_XYA  p.x p.y angle   (b2Body::GetPosition + GetAngle, F = 3; Box2D/Dynamics/b2Body.h:500-533)
 * Takes effect for the following calls; synchronises the batch's stream. */
#define B2GPU_POSE_XF 0
#define B2GPU_POSE_XYA 1
int b2gpu_batch_io_select(b2gpu_batch* batch, const int32_t* forceBodies, int32_t forceCount, const int32_t* poseBodies,
                          int32_t poseCount, int32_t poseFormat);

/* Applies `count` body commands between two steps.  Commands of one world are applied in array order (the order
 * the reference's user would have made the calls in); different worlds are independent and run in parallel.
 * The array may be in any world order (it is grouped by world on the host, stably).  A command naming a body
 * that does not exist sets B2GPU_STATUS_BAD_COMMAND on its world and is skipped.  Asynchronous on the batch's
 * stream (the command array is consumed before the call returns). */
int b2gpu_batch_body_commands(b2gpu_batch* batch, const b2gpu_body_command* commands, int64_t count);

/* Observation readback for every body in one call: out[(world * maxBodies + body) * 8] = position.xy (GetPosition),
 * angle (GetAngle), awake (1.0 / 0.0, IsAwake), linearVelocity.xy, angularVelocity, sleep time.  Replaces the per-body getters of
 * Box2D/Dynamics/b2Body.h:500-560 for an RL / game-server caller.  Synchronises. */
int b2gpu_batch_read_body_state(b2gpu_batch* batch, float* out);

/* Listener events of the last step, compacted: the b2ContactListener callbacks of every world in the reference's
 * per-world order (Begin / End / PreSolve from b2Contact::Update and b2ContactManager::Destroy, PostSolve from
 * b2Island::Report, then the TOI sub-steps' events), worlds back to back.
 *   offsets[W + 1]  offsets[w] .. offsets[w + 1] = the events of world w in `events`
 *   events[i]       one b2gpu_event
 * Call with events == NULL (or capacity 0) to get the offsets only (offsets[W] = total).  Events beyond `capacity`
 * are dropped from the copy (the offsets still describe the full list); a world whose per-step event arena overflowed
 * has B2GPU_STATUS_EVENT_OVERFLOW set and holds its first maxEvents events.  Synchronises. */
typedef struct b2gpu_event {
    int32_t type;              /* B2GPU_EVENT_* */
    int32_t fixtureA, childA;  /* b2Contact::GetFixtureA() (creation index) / GetChildIndexA() */
    int32_t fixtureB, childB;
    int32_t count;             /* Begin / End / PreSolve: manifold.pointCount; PostSolve: b2ContactImpulse::count */
    float normalImpulses[2], tangentImpulses[2]; /* PostSolve only */
} b2gpu_event;
int b2gpu_batch_read_events(b2gpu_batch* batch, int64_t* offsets, b2gpu_event* events, int64_t capacity);

/* ---- snapshot / restore (the arena snapshot is the checkpoint) ---------------------------------------
 * Every persistent field of every world (bodies, fixtures, proxies with their fat AABBs and move flags, shapes, contacts
 * with manifolds and accumulated impulses, joints with their impulses and limit states, world scalars) as one host blob;
 * per-step temporaries (constraint records, islands, events) are not part of it.  A batch restored from a snapshot
 * continues bit-identically to the batch the snapshot was taken from.  The blob is only valid for a batch created with
 * the same b2gpu_batch_desc capacities and world count.  Both calls synchronise. */
int64_t b2gpu_batch_snapshot_size(b2gpu_batch* batch);
int b2gpu_batch_snapshot_save(b2gpu_batch* batch, void* blob, int64_t size);
int b2gpu_batch_snapshot_load(b2gpu_batch* batch, const void* blob, int64_t size);

/* ---- batched read-only queries --------------------------------------------------------
 * b2World::QueryAABB / b2World::RayCast (Box2D/Dynamics/b2World.cpp:965-1020; b2DynamicTree::Query / RayCast,
 * Box2D/Collision/b2DynamicTree.h:218-303) for many (world, query) pairs in one call.  Host buffers in and out;
 * both calls synchronise.  Order queries world-fastest for coalesced arena reads. */
typedef struct b2gpu_aabb_query {
    int32_t world;
    float lower[2], upper[2];
} b2gpu_aabb_query;
/* counts[q] = number of proxies whose fat AABB overlaps query q (what b2QueryCallback::ReportFixture would be called
 * for, once per proxy); hits[q * maxHits + i] = fixture | child << 16 for the first maxHits of them in ascending
 * proxy id (the reference reports the same set in tree-traversal order). */
int b2gpu_batch_query_aabb(b2gpu_batch* batch, const b2gpu_aabb_query* queries, int64_t count, int32_t maxHits,
                           int32_t* counts, int32_t* hits);

typedef struct b2gpu_ray {
    int32_t world;
    float p1[2], p2[2];
    float maxFraction; /* b2RayCastInput::maxFraction; b2World::RayCast uses 1 */
} b2gpu_ray;
typedef struct b2gpu_ray_hit {
    int32_t fixture, child; /* fixture = -1: no hit */
    float fraction;
    float point[2], normal[2];
} b2gpu_ray_hit;
/* b2World::RayCast with the closest-hit callback (ReportFixture returns `fraction`): nearest fixture along each ray,
 * with the reference's point / normal / fraction bit for bit (exact ties in the fraction go to the lower proxy id). */
int b2gpu_batch_raycast_closest(b2gpu_batch* batch, const b2gpu_ray* rays, int64_t count, b2gpu_ray_hit* hits);

/* CUDA-event time of the last b2gpu_batch_step_n in milliseconds, and per-kernel times of a profiled step. */
int b2gpu_batch_last_step_ms(b2gpu_batch* batch, float* ms);
/* Runs `steps` steps with CUDA events around each kernel; ms[4] = collide, islands, solve, sync_pairs totals. */
int b2gpu_batch_profile_steps(b2gpu_batch* batch, float dt, int32_t velocityIterations, int32_t positionIterations,
                              int32_t steps, float* ms);
/* Same per kernel: ms[i] = total time of the i-th kernel of a step over `steps` steps, *count = kernels per step
 * (<= capacity); b2gpu_batch_kernel_name(batch, i) names them. */
int b2gpu_batch_profile_kernels(b2gpu_batch* batch, float dt, int32_t velocityIterations, int32_t positionIterations,
                                int32_t steps, float* ms, int32_t capacity, int32_t* count);
const char* b2gpu_batch_kernel_name(b2gpu_batch* batch, int32_t index);
/* Diagnostic timeline of the world groups (the step chain runs once per group of tiles, on separate streams):
 * ms[(group * steps + step) * (kernels + 1) + k] = completion time of kernel k - 1 of that step on that group's stream
 * (k = 0: start of the chain), in ms since the first chain started. */
int b2gpu_batch_trace_groups(b2gpu_batch* batch, float dt, int32_t velocityIterations, int32_t positionIterations,
                             int32_t steps, float* ms, int32_t capacity, int32_t* groups, int32_t* kernels);
/* Number of kernel launches issued by this batch so far. */
int64_t b2gpu_batch_launch_count(b2gpu_batch* batch);

const char* b2gpu_last_error(void);

/* ---- host layer -------------------------------------------------------------------------------- */

/* Capacities a batch needs for these scenes (max over scenes; contacts/events are estimates the caller
 * may raise).  Fills every max* field of desc. */
int b2gpu_scene_capacities(const b2g_scene* scenes, int32_t sceneCount, b2gpu_batch_desc* desc);

/* Replaces the b2World::CreateBody / b2Body::CreateFixture call sequence (b2World.cpp:98-122, b2Body.cpp:168-214)
 * for worlds [firstWorld, firstWorld + sceneCount): world firstWorld + i is built from scenes[i].
 * sceneCount == 1 with replicate > 1 builds `replicate` identical worlds. */
int b2gpu_batch_build_worlds(b2gpu_batch* batch, const b2g_scene* scenes, int32_t sceneCount, int32_t firstWorld,
                             int32_t replicate);

/* Replaces b2World::CreateBody + b2Body::CreateFixture calls made AFTER the first Step (b2World.cpp:98-122,
 * b2Body.cpp:168-214): spawn i creates one body from spawns[i].body with the fixtures
 * fixtures[firstFixture .. firstFixture + fixtureCount) (their `body` member is ignored; vertex indices refer to
 * `vertices`).  The host layer evaluates shapes, mass data and fat AABBs like at build time; the device appends the
 * body (new highest index = head of b2World::m_bodyList) and its fixtures, gives each proxy the most recently freed
 * proxy slot (b2DynamicTree's LIFO free list, so proxy ids — the sort key of b2BroadPhase::UpdatePairs — are the
 * reference's) or a fresh one, buffers the moves and raises e_newFixture.  Spawns of one world are applied in array
 * order.  bodyIndices[i] = the new body's index, or -1 when the world's arenas are full
 * (B2GPU_STATUS_SPAWN_OVERFLOW; raise the capacities in b2gpu_batch_desc before b2gpu_batch_create).  Synchronises. */
typedef struct b2gpu_spawn {
    int32_t world;
    int32_t firstFixture, fixtureCount;
    b2g_body_def body;
} b2gpu_spawn;
int b2gpu_batch_create_bodies(b2gpu_batch* batch, const b2gpu_spawn* spawns, int32_t count, const b2g_fixture_def* fixtures,
                              const float* vertices, int32_t* bodyIndices);
/* Device-layer primitive under it: applies pre-evaluated spawn records (layout: box2dpp_b200/csrc/b2g_host.cpp). */
int b2gpu_batch_spawn_records(b2gpu_batch* batch, const uint32_t* words, const int64_t* offsets, int32_t count,
                              int32_t* bodyIndices);

/* Replaces b2Body::CreateFixture on a body that exists already (b2Body.cpp:168-214): attach[i] adds the fixtures
 * fixtures[firstFixture .. firstFixture + fixtureCount) to body `body` of world `world`, one CreateFixture call each, in
 * order: head of the body's fixture list, proxies under the body's current transform if it is active, mass data reset
 * if the fixture has density, e_newFixture raised (the next Step starts with FindNewContacts).  fixtureIndices[i] = index
 * of the first new fixture (the handle B2GPU_CMD_DESTROY_FIXTURE / SET_FILTER take), or -1 when an arena is full
 * (B2GPU_STATUS_SPAWN_OVERFLOW) or the body does not exist (B2GPU_STATUS_BAD_COMMAND).  Synchronises. */
typedef struct b2gpu_attach {
    int32_t world, body;
    int32_t firstFixture, fixtureCount;
} b2gpu_attach;
int b2gpu_batch_create_fixtures(b2gpu_batch* batch, const b2gpu_attach* attach, int32_t count, const b2g_fixture_def* fixtures,
                                const float* vertices, int32_t* fixtureIndices);

/* Replaces b2World::CreateJoint calls made on a running world (b2World.cpp:205-265): joint i is created in world
 * worlds[i] from defs[i] (the def after Initialize(), as in b2g_scene).  The joint becomes the head of the world's and
 * of both bodies' joint lists; if collideConnected is false the contacts between its two bodies are flagged for
 * filtering, so the next Step destroys them (:255-263).  Joints of one world are created in array order.
 * jointIds[i] = the joint's id (the handle B2GPU_CMD_DESTROY_JOINT takes), or -1 when the world's joint arena is full
 * (B2GPU_STATUS_SPAWN_OVERFLOW) — size it with b2gpu_batch_desc.maxJoints.  Gear and mouse joints read the world when they
 * are constructed (b2GearJoint.cpp:47-133: joint1 / joint2 — named by joint id in enableLimit / enableMotor, as in a scene —
 * and their bodies' poses; b2MouseJoint.cpp:33-52: body B's transform): for them the call first applies the joints that
 * precede them in `defs`, reads the world's body and joint rows back and evaluates the constructor on the host.  A bad def
 * returns an error; joints that precede it in the call have been created by then.  Synchronises. */
int b2gpu_batch_create_joints(b2gpu_batch* batch, const int32_t* worlds, const b2g_joint_def* defs, int32_t count,
                              int32_t* jointIds);
/* Device-layer primitive under it: 21 words per joint (world, then the five 4-word joint columns). */
int b2gpu_batch_joint_records(b2gpu_batch* batch, const uint32_t* words, int32_t count, int32_t* jointIds);

/* Host-side shape math, exported for tests: b2PolygonShape::Set / SetAsBox / ComputeMass
 * (Box2D/Collision/Shapes/b2PolygonShape.cpp:32-245, 354-437).  out: 8 vertices, 8 normals, centroid, then
 * mass, center.xy, I (38 floats); returns the vertex count. */
int b2gpu_host_polygon_build(const b2g_fixture_def* fd, const float* vertices, float* out);

/* b2Fixture::GetMassData (Box2D/Dynamics/b2Fixture.h:347-350 -> b2Shape::ComputeMass at fd->density) of one fixture
 * description, any shape type.  out: mass, I, center.x, center.y — the payload of B2GPU_CMD_SET_DENSITY.  Host code. */
int b2gpu_host_fixture_mass_data(const b2g_fixture_def* fd, const float* vertices, int32_t vertexCount, float* out);

/* b2Body::GetMass / GetInertia-about-the-centre / GetLocalCenter of one body of a scene description, i.e.
 * b2Body::ResetMassData (Box2D/Dynamics/b2Body.cpp:283-352) over the fixtures recorded so far.  out: mass, I, center.x,
 * center.y (static and kinematic bodies: zeros).  Host code, no device: the C++ mirror answers b2Body::GetWorldCenter
 * with it while a scene is still being recorded. */
int b2gpu_host_body_mass_data(const b2g_scene* scene, int32_t body, float* out);

/* b2Rot::Set(angle) (Box2D/Common/b2Math.h:302-311) as the host layer evaluates it: the library's own
 * sinf / cosf, bit-identical to the libm the reference links (tests/test_sincos.py).  Host code, no device. */
int b2gpu_host_sincosf(const float* x, float* s, float* c, int64_t n);
/* The same values through the branch-free |x| < 120 arithmetic the position solver evaluates inline (b2g_math.cuh
 * MakeRotWideRange); exported for the bit-parity test. */
int b2gpu_host_sincosf_wide(const float* x, float* s, float* c, int64_t n);

/* ---- stateless device entry points (unit-level parity of the kernels' building blocks) -------------- */

/* The same b2Rot::Set evaluated by a CUDA kernel on `device` (host buffers in and out). */
int b2gpu_sincosf(int32_t device, const float* x, float* s, float* c, int64_t n);

/* Replaces direct calls of the narrow-phase functions (Box2D/Collision/b2Collision.h:237-300;
 * b2CollidePolygons b2CollidePolygon.cpp:119-242, b2CollideCircles / b2CollidePolygonAndCircle
 * b2CollideCircle.cpp:25-153, b2CollideEdgeAndCircle / b2CollideEdgeAndPolygon b2CollideEdge.cpp) for n
 * independent shape pairs, one CUDA thread per pair.  Flat host formats:
 *   polygon 35 floats: count, 8 vertices, 8 normals, centroid      circle 3 floats: p.xy, radius
 *   edge 10 floats: v0 v1 v2 v3, hasVertex0, hasVertex3            transform 4 floats: p.xy, q.s, q.c
 * out ints[4n]: pointCount, manifold type, id.key of point 0 / 1; floats[8n]: localNormal, localPoint,
 * points[0].localPoint, points[1].localPoint (zero when pointCount == 0). */
enum { B2GPU_COLLIDE_POLYGONS = 0, B2GPU_COLLIDE_CIRCLES = 1, B2GPU_COLLIDE_POLYGON_CIRCLE = 2,
       B2GPU_COLLIDE_EDGE_CIRCLE = 3, B2GPU_COLLIDE_EDGE_POLYGON = 4 };
int b2gpu_collide_pairs(int32_t device, int32_t kind, int32_t n, const float* shapeA, const float* xfA,
                        const float* shapeB, const float* xfB, int32_t* ints, float* floats);

#ifdef __cplusplus
}
#endif
#endif
