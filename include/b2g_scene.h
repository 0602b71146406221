/*
 * b2g_scene.h — flat, plain-C description of ONE world as the user built it:
 * the arguments of every b2World::CreateBody / b2Body::CreateFixture call, in
 * call order.  It is the common input of
 *   - the product's host layer (box2dpp_b200/csrc/b2g_host.cpp), which turns it
 *     into world-interleaved device arenas, and
 *   - the test oracles under oracle/ (the vanilla-library harness and the C
 *     restatement), which rebuild the same world on the CPU.
 *
 * Reference interfaces this mirrors (paths relative to /root/reference):
 *   b2BodyDef      Box2D/Dynamics/b2Body.h:53-125
 *   b2FixtureDef   Box2D/Dynamics/b2Fixture.h:59-95,  b2Filter :35-57
 *   b2CircleShape / b2EdgeShape / b2PolygonShape / b2ChainShape
 *                  Box2D/Collision/Shapes/ (all four headers)
 *   b2World(gravity) + Set* flags   Box2D/Dynamics/b2World.h:48-232
 */
#ifndef B2G_SCENE_H
#define B2G_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* b2BodyType — Box2D/Dynamics/b2Body.h:41-49 */
enum { B2G_STATIC_BODY = 0, B2G_KINEMATIC_BODY = 1, B2G_DYNAMIC_BODY = 2 };

/* b2Shape::Type — Box2D/Collision/Shapes/b2Shape.h:47-54 */
enum { B2G_SHAPE_CIRCLE = 0, B2G_SHAPE_EDGE = 1, B2G_SHAPE_POLYGON = 2, B2G_SHAPE_CHAIN = 3 };

/* How a shape was built by the user (which shape-class method was called). */
enum {
    B2G_POLY_BOX        = 0, /* SetAsBox(hx, hy)                       b2PolygonShape.cpp:32-44  */
    B2G_POLY_BOX_OFFSET = 1, /* SetAsBox(hx, hy, center, angle)        b2PolygonShape.cpp:46-71  */
    B2G_POLY_SET        = 2, /* Set(points, count) (welding + hull)    b2PolygonShape.cpp:122-245 */
    B2G_CHAIN_OPEN      = 0, /* CreateChain(vertices, count)           b2ChainShape.cpp:63-79    */
    B2G_CHAIN_LOOP      = 1  /* CreateLoop(vertices, count)            b2ChainShape.cpp:41-61    */
};

/* body-def flag bits */
enum {
    B2G_BD_ALLOW_SLEEP    = 1,
    B2G_BD_AWAKE          = 2,
    B2G_BD_FIXED_ROTATION = 4,
    B2G_BD_BULLET         = 8,
    B2G_BD_ACTIVE         = 16,
    B2G_BD_DEFAULT_FLAGS  = 1 | 2 | 16
};

/* world flag bits (b2World::SetAllowSleeping / SetWarmStarting / SetContinuousPhysics /
 * SetAutoClearForces; b2World.h:146-180,214-217; defaults b2World.cpp:39-52) */
enum {
    B2G_WORLD_ALLOW_SLEEP   = 1,
    B2G_WORLD_WARM_STARTING = 2,
    B2G_WORLD_CONTINUOUS    = 4, /* TOI: not implemented on the device path */
    B2G_WORLD_CLEAR_FORCES  = 8,
    B2G_WORLD_DEFAULT_FLAGS = 1 | 2 | 4 | 8
};

typedef struct b2g_body_def {
    int32_t type;            /* B2G_*_BODY */
    int32_t flags;           /* B2G_BD_* */
    float position[2];
    float angle;
    float linearVelocity[2];
    float angularVelocity;
    float linearDamping;
    float angularDamping;
    float gravityScale;
} b2g_body_def;

typedef struct b2g_fixture_def {
    int32_t body;            /* index into bodies[] (creation order) */
    int32_t shapeType;       /* B2G_SHAPE_* */
    int32_t shapeKind;       /* B2G_POLY_* or B2G_CHAIN_* */
    int32_t isSensor;
    float friction;          /* default 0.2  */
    float restitution;       /* default 0    */
    float density;           /* default 0    */
    uint16_t categoryBits;   /* default 1      */
    uint16_t maskBits;       /* default 0xFFFF */
    int16_t groupIndex;      /* default 0      */
    int16_t pad0;
    float radius;            /* circle: m_radius. others: ignored (b2_polygonRadius) */
    /* circle: p = m_p.  box: p = (hx, hy), center, angle. */
    float p[2];
    float center[2];
    float angle;
    /* polygon Set / chain / edge: vertices[firstVertex .. firstVertex+vertexCount)
       edge: 2 vertices (v1, v2).  */
    int32_t firstVertex;
    int32_t vertexCount;
    /* edge: ghost vertices v0 / v3; chain: prev / next vertex (SetPrevVertex/SetNextVertex) */
    int32_t hasVertex0, hasVertex3;
    float vertex0[2], vertex3[2];
} b2g_fixture_def;

/* Joint types: the values of b2JointType (Box2D/Dynamics/Joints/b2Joint.h:31-45). */
enum { B2G_JOINT_REVOLUTE = 1, B2G_JOINT_PRISMATIC = 2, B2G_JOINT_DISTANCE = 3, B2G_JOINT_PULLEY = 4, B2G_JOINT_MOUSE = 5, B2G_JOINT_GEAR = 6,
       B2G_JOINT_WHEEL = 7, B2G_JOINT_WELD = 8, B2G_JOINT_FRICTION = 9, B2G_JOINT_ROPE = 10, B2G_JOINT_MOTOR = 11 };

/* The arguments of b2World::CreateJoint: the members of b2RevoluteJointDef (b2RevoluteJoint.h:35-86),
 * b2PrismaticJointDef (b2PrismaticJoint.h:33-87), b2DistanceJointDef (b2DistanceJoint.h:32-67), b2WheelJointDef
 * (b2WheelJoint.h:33-73), b2WeldJointDef (b2WeldJoint.h:30-62), b2RopeJointDef, b2FrictionJointDef, b2MotorJointDef,
 * b2PulleyJointDef, b2MouseJointDef and b2GearJointDef (all eleven b2JointType values), already in the bodies' local
 * frames (what the defs' Initialize() helpers compute).  One record serves every type; the comments on the fields say
 * which def member each type keeps where. */
typedef struct b2g_joint_def {
    int32_t type;             /* B2G_JOINT_* */
    int32_t bodyA, bodyB;     /* indices into bodies[] */
    int32_t collideConnected; /* b2JointDef::collideConnected, default false */
    float localAnchorA[2], localAnchorB[2];
    float localAxisA[2];      /* prismatic (normalised by the joint's constructor), wheel (used as given) */
    float referenceAngle;     /* revolute, prismatic, weld */
    int32_t enableLimit, enableMotor;
    float lower, upper;       /* revolute: lowerAngle / upperAngle; prismatic: lowerTranslation / upperTranslation */
    float maxMotor;           /* revolute, wheel: maxMotorTorque; prismatic: maxMotorForce; friction: maxTorque */
    float motorSpeed;
    float length;             /* distance: length; rope: maxLength; friction, motor: maxForce */
                              /* pulley: localAxisA = groundAnchorA, (lower, upper) = groundAnchorB, length = lengthA,
                                 maxMotor = lengthB, motorSpeed = ratio */
                              /* motor: localAnchorA = linearOffset, referenceAngle = angularOffset, maxMotor = maxTorque,
                                 frequencyHz = correctionFactor */
                              /* mouse (b2MouseJointDef, b2MouseJoint.h:29-55): localAnchorA = target (world point),
                                 length = maxForce, frequencyHz, dampingRatio; bodyA is unused by the solver */
                              /* gear (b2GearJointDef, b2GearJoint.h:27-47): enableLimit = index of joint1, enableMotor =
                                 index of joint2 in joints[] (revolute or prismatic, created earlier), motorSpeed = ratio;
                                 the bodies come from those joints, as in b2GearJoint's constructor */
    float frequencyHz, dampingRatio; /* distance, wheel, weld, mouse */
} b2g_joint_def;

typedef struct b2g_scene {
    float gravity[2];
    int32_t worldFlags;      /* B2G_WORLD_* */
    int32_t bodyCount;
    int32_t fixtureCount;
    int32_t vertexCount;
    const b2g_body_def* bodies;       /* creation order */
    const b2g_fixture_def* fixtures;  /* creation order (global, across bodies) */
    const float* vertices;            /* xy pairs */
    const b2g_joint_def* joints;      /* b2World::CreateJoint calls in call order (may be NULL when jointCount == 0) */
    int32_t jointCount;
    int32_t pad0;
} b2g_scene;

#ifdef __cplusplus
}
#endif
#endif
