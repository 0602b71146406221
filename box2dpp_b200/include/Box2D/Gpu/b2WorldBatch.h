// b2WorldBatch.h — the reference's scene-construction and stepping API over the B200 batch engine.
//
// box2dpp's boundary is its C++ public API (there is no FFI layer in the reference).  This header keeps that
// API's spelling — namespace box2d, b2Vec<float, 2>, b2BodyDef / b2FixtureDef / the four shape classes,
// b2World::CreateBody, b2Body::CreateFixture, b2World::Step — and adds the batch axis: a b2WorldBatch owns N
// independent b2World objects that are built through the same calls and stepped together on one GPU by
// b2gpu_batch_step (include/b2gpu.h), the replacement of b2World::Step (Box2D/Dynamics/b2World.cpp:888-954).
// A b2World constructed on its own is a batch of one, so HelloWorld/HelloWorld.cpp:36-101 compiles against this
// header unchanged apart from the include line (tests/cpp/hello_world_batch.cpp; continuous physics — b2World::SolveTOI —
// is on by default, like the reference's).
//
// Reference interfaces mirrored (paths under /root/reference/Box2D):
//   b2Vec<T, N>, b2VecX / b2VecY              Common/b2Vec.h:17-72
//   b2BodyType, b2BodyDef, b2Body getters     Dynamics/b2Body.h:41-125, 163-197, 500-533, 738-790
//   b2Filter, b2FixtureDef, b2Fixture         Dynamics/b2Fixture.h:35-95
//   b2Shape and the four shapes               Collision/Shapes/b2Shape.h:44-128, b2CircleShape.h, b2EdgeShape.h:35,
//                                             b2PolygonShape.h:48-60, b2ChainShape.h:48-61
//   b2World                                   Dynamics/b2World.h:48-232
//   b2ContactListener, b2ContactImpulse       Dynamics/b2WorldCallbacks.h:74-136
//   b2QueryCallback, b2RayCastCallback        Dynamics/b2WorldCallbacks.h:122-153
//   the eleven b2*JointDef, b2Joint           Dynamics/Joints/*.h
//
// Differences a caller must know:
//   * The scene is recorded on the host until the first Step (or the first call that needs the device); from then on
//     CreateBody / CreateFixture / CreateJoint / DestroyBody / DestroyFixture / DestroyJoint and the b2Body / b2Fixture
//     setters below act on the running worlds with the reference's semantics (arena headroom: SetSpawnHeadroom,
//     SetJointHeadroom).
//   * Shapes record HOW they were built (SetAsBox / Set / CreateChain ...); hulls, normals, mass data and fat
//     AABBs are computed by the host layer of libb2gpu.so with the reference's arithmetic when the batch is built.
//   * Listener callbacks are delivered after Step from the device's readback buffer, in the reference's order,
//     as b2ContactEvent records; they cannot mutate the contact mid-step (PreSolve is read-only).
//   * Errors: misuse that the reference asserts on throws std::runtime_error carrying b2gpu_last_error().
#ifndef B2_WORLD_BATCH_H
#define B2_WORLD_BATCH_H

#include <array>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b2gpu.h"

namespace box2d {

// ---- b2Vec (Common/b2Vec.h) ------------------------------------------------------------------------
template <class T, std::size_t N>
struct b2Vec {
    std::array<T, N> v;
    T& operator[](std::size_t i) { return v[i]; }
    const T& operator[](std::size_t i) const { return v[i]; }
};
constexpr std::size_t b2VecX = 0, b2VecY = 1;
using b2Vec2f = b2Vec<float, 2>;

enum class b2BodyType { STATIC_BODY = 0, KINEMATIC_BODY, DYNAMIC_BODY };

struct b2BodyDef {
    b2BodyType type = b2BodyType::STATIC_BODY;
    b2Vec<float, 2> position{{0.0f, 0.0f}};
    float angle = 0.0f;
    b2Vec<float, 2> linearVelocity{{0.0f, 0.0f}};
    float angularVelocity = 0.0f;
    float linearDamping = 0.0f;
    float angularDamping = 0.0f;
    bool allowSleep = true;
    bool awake = true;
    bool fixedRotation = false;
    bool bullet = false;
    bool active = true;
    void* userData = nullptr;
    float gravityScale = 1.0f;
};

struct b2Filter {
    uint16_t categoryBits = 0x0001;
    uint16_t maskBits = 0xFFFF;
    int16_t groupIndex = 0;
};

// ---- shapes ----------------------------------------------------------------------------------------
class b2Shape {
public:
    enum Type { e_circle = 0, e_edge = 1, e_polygon = 2, e_chain = 3, e_typeCount = 4 };
    virtual ~b2Shape() {}
    Type GetType() const { return m_type; }
    float GetRadius() const { return m_radius; }
    void SetRadius(float radius) { m_radius = radius; }

protected:
    friend class b2Body;
    explicit b2Shape(Type t, float r) : m_type(t), m_radius(r) {}
    // appends this shape's construction record to a fixture def + vertex pool (include/b2g_scene.h)
    virtual void Describe(b2g_fixture_def& fd, std::vector<float>& vertices) const = 0;
    Type m_type;
    float m_radius;
};

class b2CircleShape : public b2Shape {
public:
    b2CircleShape() : b2Shape(e_circle, 0.0f) {}
    b2Vec<float, 2> m_p{{0.0f, 0.0f}};

protected:
    void Describe(b2g_fixture_def& fd, std::vector<float>&) const override {
        fd.shapeType = B2G_SHAPE_CIRCLE;
        fd.radius = m_radius;
        fd.p[0] = m_p[0];
        fd.p[1] = m_p[1];
    }
};

class b2EdgeShape : public b2Shape {
public:
    b2EdgeShape() : b2Shape(e_edge, 0.01f) {}
    void Set(b2Vec<float, 2> v1, b2Vec<float, 2> v2) {
        m_vertex1 = v1;
        m_vertex2 = v2;
        m_hasVertex0 = m_hasVertex3 = false;
    }
    b2Vec<float, 2> m_vertex0{{0, 0}}, m_vertex1{{0, 0}}, m_vertex2{{0, 0}}, m_vertex3{{0, 0}};
    bool m_hasVertex0 = false, m_hasVertex3 = false;

protected:
    void Describe(b2g_fixture_def& fd, std::vector<float>& vertices) const override {
        fd.shapeType = B2G_SHAPE_EDGE;
        fd.firstVertex = (int32_t)(vertices.size() / 2);
        fd.vertexCount = 2;
        vertices.insert(vertices.end(), {m_vertex1[0], m_vertex1[1], m_vertex2[0], m_vertex2[1]});
        fd.hasVertex0 = m_hasVertex0;
        fd.hasVertex3 = m_hasVertex3;
        fd.vertex0[0] = m_vertex0[0]; fd.vertex0[1] = m_vertex0[1];
        fd.vertex3[0] = m_vertex3[0]; fd.vertex3[1] = m_vertex3[1];
    }
};

class b2PolygonShape : public b2Shape {
public:
    b2PolygonShape() : b2Shape(e_polygon, 0.01f) {}
    void SetAsBox(float hx, float hy) {
        m_kind = B2G_POLY_BOX;
        m_hx = hx;
        m_hy = hy;
    }
    void SetAsBox(float hx, float hy, const b2Vec<float, 2>& center, float angle) {
        m_kind = B2G_POLY_BOX_OFFSET;
        m_hx = hx;
        m_hy = hy;
        m_center = center;
        m_angle = angle;
    }
    void Set(const b2Vec<float, 2>* points, int32_t count) {
        if (count < 3 || count > 8) throw std::runtime_error("b2PolygonShape::Set: 3 <= count <= b2_maxPolygonVertices");
        m_kind = B2G_POLY_SET;
        m_points.assign(points, points + count);
    }

protected:
    void Describe(b2g_fixture_def& fd, std::vector<float>& vertices) const override {
        fd.shapeType = B2G_SHAPE_POLYGON;
        fd.shapeKind = m_kind;
        fd.p[0] = m_hx; fd.p[1] = m_hy;
        fd.center[0] = m_center[0]; fd.center[1] = m_center[1];
        fd.angle = m_angle;
        if (m_kind == B2G_POLY_SET) {
            fd.firstVertex = (int32_t)(vertices.size() / 2);
            fd.vertexCount = (int32_t)m_points.size();
            for (const auto& p : m_points) { vertices.push_back(p[0]); vertices.push_back(p[1]); }
        }
    }
    int m_kind = B2G_POLY_BOX;
    float m_hx = 0.0f, m_hy = 0.0f, m_angle = 0.0f;
    b2Vec<float, 2> m_center{{0, 0}};
    std::vector<b2Vec<float, 2>> m_points;
};

class b2ChainShape : public b2Shape {
public:
    b2ChainShape() : b2Shape(e_chain, 0.01f) {}
    void CreateLoop(const b2Vec<float, 2>* vertices, int32_t count) { m_loop = true; m_points.assign(vertices, vertices + count); }
    void CreateChain(const b2Vec<float, 2>* vertices, int32_t count) { m_loop = false; m_points.assign(vertices, vertices + count); }
    void SetPrevVertex(const b2Vec<float, 2>& v) { m_prev = v; m_hasPrev = true; }
    void SetNextVertex(const b2Vec<float, 2>& v) { m_next = v; m_hasNext = true; }

protected:
    void Describe(b2g_fixture_def& fd, std::vector<float>& vertices) const override {
        fd.shapeType = B2G_SHAPE_CHAIN;
        fd.shapeKind = m_loop ? B2G_CHAIN_LOOP : B2G_CHAIN_OPEN;
        fd.firstVertex = (int32_t)(vertices.size() / 2);
        fd.vertexCount = (int32_t)m_points.size();
        for (const auto& p : m_points) { vertices.push_back(p[0]); vertices.push_back(p[1]); }
        fd.hasVertex0 = m_hasPrev; fd.hasVertex3 = m_hasNext;
        fd.vertex0[0] = m_prev[0]; fd.vertex0[1] = m_prev[1];
        fd.vertex3[0] = m_next[0]; fd.vertex3[1] = m_next[1];
    }
    bool m_loop = false, m_hasPrev = false, m_hasNext = false;
    b2Vec<float, 2> m_prev{{0, 0}}, m_next{{0, 0}};
    std::vector<b2Vec<float, 2>> m_points;
};

struct b2FixtureDef {
    const b2Shape* shape = nullptr;
    void* userData = nullptr;
    float friction = 0.2f;
    float restitution = 0.0f;
    float density = 0.0f;
    bool isSensor = false;
    b2Filter filter;
};

class b2Body;
class b2World;
class b2WorldBatch;

// ---- joints (Dynamics/Joints/b2Joint.h:31-101, b2RevoluteJoint.h:35-86, b2PrismaticJoint.h:33-87, b2DistanceJoint.h:32-67,
// b2WheelJoint.h:33-73, b2WeldJoint.h:30-62)
enum class b2JointType {
    UNKNOWN_JOINT, REVOLUTE_JOINT, PRISMATIC_JOINT, DISTANCE_JOINT, PULLEY_JOINT, MOUSE_JOINT, GEAR_JOINT, WHEEL_JOINT,
    WELD_JOINT, FRICTION_JOINT, ROPE_JOINT, MOTOR_JOINT
};
struct b2JointDef {
    b2JointType type = b2JointType::UNKNOWN_JOINT;
    void* userData = nullptr;
    b2Body* bodyA = nullptr;
    b2Body* bodyB = nullptr;
    bool collideConnected = false;
};
struct b2RevoluteJointDef : public b2JointDef {
    b2RevoluteJointDef() { type = b2JointType::REVOLUTE_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchor);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}};
    float referenceAngle = 0.0f;
    bool enableLimit = false;
    float lowerAngle = 0.0f, upperAngle = 0.0f;
    bool enableMotor = false;
    float motorSpeed = 0.0f, maxMotorTorque = 0.0f;
};
struct b2PrismaticJointDef : public b2JointDef {
    b2PrismaticJointDef() { type = b2JointType::PRISMATIC_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchor, const b2Vec<float, 2>& axis);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}}, localAxisA{{1.0f, 0.0f}};
    float referenceAngle = 0.0f;
    bool enableLimit = false;
    float lowerTranslation = 0.0f, upperTranslation = 0.0f;
    bool enableMotor = false;
    float maxMotorForce = 0.0f, motorSpeed = 0.0f;
};
struct b2DistanceJointDef : public b2JointDef {
    b2DistanceJointDef() { type = b2JointType::DISTANCE_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchorA, const b2Vec<float, 2>& anchorB);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}};
    float length = 1.0f, frequencyHz = 0.0f, dampingRatio = 0.0f;
};
struct b2WheelJointDef : public b2JointDef {
    b2WheelJointDef() { type = b2JointType::WHEEL_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchor, const b2Vec<float, 2>& axis);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}}, localAxisA{{1.0f, 0.0f}};
    bool enableMotor = false;
    float maxMotorTorque = 0.0f, motorSpeed = 0.0f;
    float frequencyHz = 2.0f, dampingRatio = 0.7f;
};
struct b2WeldJointDef : public b2JointDef {
    b2WeldJointDef() { type = b2JointType::WELD_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchor);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}};
    float referenceAngle = 0.0f, frequencyHz = 0.0f, dampingRatio = 0.0f;
};
struct b2RopeJointDef : public b2JointDef {  // Dynamics/Joints/b2RopeJoint.h:28-52
    b2RopeJointDef() { type = b2JointType::ROPE_JOINT; }
    b2Vec<float, 2> localAnchorA{{-1.0f, 0.0f}}, localAnchorB{{1.0f, 0.0f}};
    float maxLength = 0.0f;
};
struct b2FrictionJointDef : public b2JointDef {  // Dynamics/Joints/b2FrictionJoint.h:27-56
    b2FrictionJointDef() { type = b2JointType::FRICTION_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& anchor);
    b2Vec<float, 2> localAnchorA{{0.0f, 0.0f}}, localAnchorB{{0.0f, 0.0f}};
    float maxForce = 0.0f, maxTorque = 0.0f;
};
struct b2PulleyJointDef : public b2JointDef {  // Dynamics/Joints/b2PulleyJoint.h:29-72
    b2PulleyJointDef() { type = b2JointType::PULLEY_JOINT; collideConnected = true; }
    void Initialize(b2Body* bodyA, b2Body* bodyB, const b2Vec<float, 2>& groundAnchorA, const b2Vec<float, 2>& groundAnchorB,
                    const b2Vec<float, 2>& anchorA, const b2Vec<float, 2>& anchorB, float ratio);
    b2Vec<float, 2> groundAnchorA{{-1.0f, 1.0f}}, groundAnchorB{{1.0f, 1.0f}}, localAnchorA{{-1.0f, 0.0f}}, localAnchorB{{1.0f, 0.0f}};
    float lengthA = 0.0f, lengthB = 0.0f, ratio = 1.0f;
};
struct b2MouseJointDef : public b2JointDef {  // Dynamics/Joints/b2MouseJoint.h:29-55
    b2MouseJointDef() { type = b2JointType::MOUSE_JOINT; }
    b2Vec<float, 2> target{{0.0f, 0.0f}};
    float maxForce = 0.0f, frequencyHz = 5.0f, dampingRatio = 0.7f;
};
class b2Joint;
struct b2GearJointDef : public b2JointDef {  // Dynamics/Joints/b2GearJoint.h:27-47
    b2GearJointDef() { type = b2JointType::GEAR_JOINT; }
    b2Joint* joint1 = nullptr;  // revolute or prismatic
    b2Joint* joint2 = nullptr;
    float ratio = 1.0f;
};
struct b2MotorJointDef : public b2JointDef {  // Dynamics/Joints/b2MotorJoint.h:27-57
    b2MotorJointDef() { type = b2JointType::MOTOR_JOINT; }
    void Initialize(b2Body* bodyA, b2Body* bodyB);
    b2Vec<float, 2> linearOffset{{0.0f, 0.0f}};
    float angularOffset = 0.0f, maxForce = 1.0f, maxTorque = 1.0f, correctionFactor = 0.3f;
};
class b2Joint {
public:
    b2JointType GetType() const { return m_type; }
    b2Body* GetBodyA() const { return m_bodyA; }
    b2Body* GetBodyB() const { return m_bodyB; }
    void* GetUserData() const { return m_userData; }
    int32_t GetIndex() const { return m_index; }  // creation order within the world
    // b2MouseJoint::SetTarget (Dynamics/Joints/b2MouseJoint.cpp:54-61); ignored for the other joint types.  Queued like
    // the b2Body edits and applied before the next Step.
    void SetTarget(const b2Vec<float, 2>& target);

private:
    friend class b2World;
    b2JointType m_type = b2JointType::UNKNOWN_JOINT;
    b2Body* m_bodyA = nullptr;
    b2Body* m_bodyB = nullptr;
    void* m_userData = nullptr;
    int32_t m_index = 0;
};

class b2Fixture {
public:
    b2Body* GetBody() const { return m_body; }
    void* GetUserData() const { return m_userData; }
    b2Shape::Type GetType() const { return m_type; }
    int32_t GetIndex() const { return m_index; }  // creation order within the world
    // b2Fixture::SetFilterData (b2Fixture.cpp:185-215): the fixture's contacts are re-filtered at the next Step
    void SetFilterData(const b2Filter& filter);
    // b2Fixture::SetSensor (b2Fixture.cpp:222-229), SetFriction / SetRestitution (b2Fixture.h:330-345)
    void SetSensor(bool sensor);
    void SetFriction(float friction) { m_friction = friction; QueueMaterial(); }
    void SetRestitution(float restitution) { m_restitution = restitution; QueueMaterial(); }
    // b2Fixture::SetDensity / GetDensity (b2Fixture.h:313-323): like the reference's, the body's mass only follows at the next
    // b2Body::ResetMassData().  The fixture's b2MassData under the new density is computed on the host (b2Shape::ComputeMass,
    // b2gpu_host_fixture_mass_data) and travels in the command.
    void SetDensity(float density);
    float GetDensity() const { return m_def.density; }
    // the getters beside them (b2Fixture.h:120-190): host copies of what the last setter / the definition said
    float GetFriction() const { return m_friction; }
    float GetRestitution() const { return m_restitution; }
    bool IsSensor() const { return m_def.isSensor != 0; }
    b2Filter GetFilterData() const {
        b2Filter f;
        f.categoryBits = (uint16_t)m_def.categoryBits; f.maskBits = (uint16_t)m_def.maskBits; f.groupIndex = (int16_t)m_def.groupIndex;
        return f;
    }

private:
    friend class b2Body;
    b2Body* m_body = nullptr;
    void* m_userData = nullptr;
    b2Shape::Type m_type = b2Shape::e_circle;
    int32_t m_index = 0;
    void QueueMaterial();
    float m_friction = 0.2f, m_restitution = 0.0f;  // host copies: the device command carries both
    b2g_fixture_def m_def;                          // the fixture as it was described, firstVertex = 0 into m_defVertices
    std::vector<float> m_defVertices;
};
struct b2MassData {  // Collision/Shapes/b2Shape.h:27-38
    float mass = 0.0f;
    b2Vec<float, 2> center{{0.0f, 0.0f}};
    float I = 0.0f;
};

// ---- listener: b2ContactListener over the readback buffer -------------------------------------------
struct b2ContactImpulse {  // Dynamics/b2WorldCallbacks.h:74-79
    float normalImpulses[2];
    float tangentImpulses[2];
    int32_t count;
};
struct b2ContactEvent {  // what a b2Contact* tells a listener, by value
    b2World* world;
    b2Fixture* fixtureA;
    b2Fixture* fixtureB;
    int32_t childIndexA, childIndexB;
    int32_t pointCount;
};
// ---- queries: b2World::QueryAABB / RayCast over the batched device calls ----------------------------
struct b2AABB {  // Collision/b2Collision.h:165-223
    b2Vec<float, 2> lowerBound{{0.0f, 0.0f}};
    b2Vec<float, 2> upperBound{{0.0f, 0.0f}};
};
class b2QueryCallback {  // Dynamics/b2WorldCallbacks.h:122-131
public:
    virtual ~b2QueryCallback() {}
    virtual bool ReportFixture(b2Fixture* fixture) = 0;  // false ends the query
};
class b2RayCastCallback {  // Dynamics/b2WorldCallbacks.h:133-153
public:
    virtual ~b2RayCastCallback() {}
    virtual float ReportFixture(b2Fixture* fixture, const b2Vec<float, 2>& point, const b2Vec<float, 2>& normal, float fraction) = 0;
};
class b2ContactListener {
public:
    virtual ~b2ContactListener() {}
    virtual void BeginContact(const b2ContactEvent&) {}
    virtual void EndContact(const b2ContactEvent&) {}
    virtual void PreSolve(const b2ContactEvent&) {}
    virtual void PostSolve(const b2ContactEvent&, const b2ContactImpulse&) {}
};

class b2Body {
public:
    b2Fixture* CreateFixture(const b2FixtureDef* def);
    b2Fixture* CreateFixture(const b2Shape* shape, float density) {
        b2FixtureDef def;
        def.shape = shape;
        def.density = density;
        return CreateFixture(&def);
    }
    const b2Vec<float, 2>& GetPosition() const;
    float GetAngle() const;
    const b2Vec<float, 2>& GetWorldCenter() const;
    const b2Vec<float, 2>& GetLinearVelocity() const;
    float GetAngularVelocity() const;
    bool IsAwake() const;
    float GetMass() const;  // b2Body::GetMass (b2Body.h:530-533): m_mass as the last ResetMassData / SetMassData left it
    b2BodyType GetType() const { return m_type; }
    void* GetUserData() const { return m_userData; }
    b2World* GetWorld() const { return m_world; }
    int32_t GetIndex() const { return m_index; }  // creation order within the world
    // b2Body::GetLocalPoint / GetLocalVector (b2Body.h:567-575) — before the first Step they are evaluated from the
    // body definition, which is what the joint defs' Initialize() helpers need
    b2Vec<float, 2> GetLocalPoint(const b2Vec<float, 2>& worldPoint) const;
    b2Vec<float, 2> GetLocalVector(const b2Vec<float, 2>& worldVector) const;
    // b2Body::GetWorldPoint / GetWorldVector / GetLinearVelocityFromWorldPoint / FromLocalPoint (b2Body.h:557-587)
    b2Vec<float, 2> GetWorldPoint(const b2Vec<float, 2>& localPoint) const;
    b2Vec<float, 2> GetWorldVector(const b2Vec<float, 2>& localVector) const;
    b2Vec<float, 2> GetLinearVelocityFromWorldPoint(const b2Vec<float, 2>& worldPoint) const;
    b2Vec<float, 2> GetLinearVelocityFromLocalPoint(const b2Vec<float, 2>& localPoint) const {
        return GetLinearVelocityFromWorldPoint(GetWorldPoint(localPoint));
    }
    // flag getters (b2Body.h:634-690): the device's flag word after the first Step, the definition before
    bool IsBullet() const { return HasFlag(1 << 5, B2G_BD_BULLET); }
    bool IsSleepingAllowed() const { return HasFlag(1 << 4, B2G_BD_ALLOW_SLEEP); }
    bool IsFixedRotation() const { return HasFlag(1 << 6, B2G_BD_FIXED_ROTATION); }
    bool IsActive() const { return HasFlag(1 << 7, B2G_BD_ACTIVE); }
    float GetLinearDamping() const { return m_linearDamping; }
    float GetAngularDamping() const { return m_angularDamping; }
    // Between-step edits (Dynamics/b2Body.h:738-833, 408-467, 634-653; b2Body.cpp:418-440).  Each call is queued as a
    // b2gpu_body_command in call order and applied on the device before the next Step / getter, with the
    // reference's semantics (type checks, wake rules, force only accumulates on awake bodies, SetTransform
    // re-synchronises the body's proxies).
    void ApplyForce(const b2Vec<float, 2>& force, const b2Vec<float, 2>& point, bool wake);
    void ApplyForceToCenter(const b2Vec<float, 2>& force, bool wake);
    void ApplyTorque(float torque, bool wake);
    void ApplyLinearImpulse(const b2Vec<float, 2>& impulse, const b2Vec<float, 2>& point, bool wake);
    void ApplyAngularImpulse(float impulse, bool wake);
    void SetLinearVelocity(const b2Vec<float, 2>& v);
    void SetAngularVelocity(float omega);
    void SetTransform(const b2Vec<float, 2>& position, float angle);
    void SetAwake(bool flag);
    // Edits that change what a body is (b2Body.cpp:115-166, 216-281, 283-394, 455-510): applied on the device before the
    // next Step; calling one while the scene is still open builds the batch first.
    void SetType(b2BodyType type);
    void SetActive(bool flag);
    void SetMassData(const b2MassData* data);
    void ResetMassData();
    void DestroyFixture(b2Fixture* fixture);  // the b2Fixture object stays allocated (its index is reserved)
    // plain setters (b2Body.h:590-690, b2Body.cpp:512-534)
    void SetLinearDamping(float damping) { m_linearDamping = damping; QueueDamping(); }
    void SetAngularDamping(float damping) { m_angularDamping = damping; QueueDamping(); }
    void SetGravityScale(float scale);
    void SetBullet(bool flag);
    void SetSleepingAllowed(bool flag);
    void SetFixedRotation(bool flag);

private:
    friend class b2World;
    friend class b2WorldBatch;
    friend class b2Joint;
    friend class b2Fixture;
    void SetMouseTarget(const b2Vec<float, 2>& target);  // b2Joint::SetTarget of the mouse joint attached to this body
    bool HasFlag(int deviceBit, int defBit) const;
    void QueueDamping();
    float m_linearDamping = 0.0f, m_angularDamping = 0.0f;  // host copies: the device command carries both
    // While the scene is still being recorded (no Step yet) edits are queued and getters are answered on the host from
    // this shadow of the body's definition, so a `body->SetLinearVelocity(..)` or `jd.Initialize(a, b, a->GetWorldCenter())`
    // right after CreateBody neither creates the device batch nor freezes the scene.
    struct Shadow {
        float p[2] = {0.0f, 0.0f}, angle = 0.0f, v[2] = {0.0f, 0.0f}, w = 0.0f;
        bool awake = true;
        bool exact = true;      // false after an edit whose effect needs the device (impulses): getters then build the batch
        bool destroyed = false;
        int32_t fixtures = 0;       // fixtures given to this body so far
        int32_t massFixtures = -1;  // fixture count the cached local centre was computed for
        float localCenter[2] = {0.0f, 0.0f};
    };
    mutable Shadow m_shadow;
    b2World* m_world = nullptr;
    int32_t m_index = 0;
    b2BodyType m_type = b2BodyType::STATIC_BODY;
    void* m_userData = nullptr;
    mutable b2Vec<float, 2> m_tmp[3];
};

class b2World {
public:
    // A world on its own is a batch of one (HelloWorld/HelloWorld.cpp:39).
    explicit b2World(const b2Vec<float, 2>& gravity);
    ~b2World();
    b2World(const b2World&) = delete;
    b2World& operator=(const b2World&) = delete;

    b2Body* CreateBody(const b2BodyDef* def);
    // b2World::CreateJoint (b2World.cpp:211-270) for all eleven joint types, before the first Step or on the running world
    // (gear and mouse joints read the world when they are constructed: b2gpu_batch_create_joints evaluates that on the
    // world's current state)
    b2Joint* CreateJoint(const b2JointDef* def);
    // b2World::DestroyJoint (b2World.cpp:267-360); the b2Joint object stays allocated (its index is reserved)
    void DestroyJoint(b2Joint* joint);
    // b2World::DestroyBody (b2World.cpp:124-203): queued like the b2Body edits; the body's joints, contacts and proxies go
    // with it on the device.  The b2Body object stays allocated (its index is reserved) but must not be used again.
    // EndContact events of the destruction are in the readback buffer (b2gpu_batch_read_events) until the next Step.
    void DestroyBody(b2Body* body);
    void Step(float timeStep, int32_t velocityIterations, int32_t positionIterations);
    void SetContactListener(b2ContactListener* listener);
    void SetAllowSleeping(bool flag) { SetFlag(B2G_WORLD_ALLOW_SLEEP, flag); }
    bool GetAllowSleeping() const { return (m_flags & B2G_WORLD_ALLOW_SLEEP) != 0; }
    void SetWarmStarting(bool flag) { SetFlag(B2G_WORLD_WARM_STARTING, flag); }
    bool GetWarmStarting() const { return (m_flags & B2G_WORLD_WARM_STARTING) != 0; }
    void SetContinuousPhysics(bool flag) { SetFlag(B2G_WORLD_CONTINUOUS, flag); }
    bool GetContinuousPhysics() const { return (m_flags & B2G_WORLD_CONTINUOUS) != 0; }
    void SetAutoClearForces(bool flag) { SetFlag(B2G_WORLD_CLEAR_FORCES, flag); }
    bool GetAutoClearForces() const { return (m_flags & B2G_WORLD_CLEAR_FORCES) != 0; }
    b2Vec<float, 2> GetGravity() const { return m_gravity; }
    // b2World::QueryAABB (b2World.cpp:965-972): ReportFixture once per proxy whose fat AABB overlaps `aabb` — the
    // reference's set; the order is ascending proxy id where the reference reports in tree-traversal order.
    void QueryAABB(b2QueryCallback* callback, const b2AABB& aabb);
    // b2World::RayCast (b2World.cpp:1003-1020) for the closest-hit callback (the one that returns `fraction`): the nearest
    // fixture along the ray is reported ONCE, with the reference's point / normal / fraction.  The reference also reports
    // farther fixtures it meets on its way through the tree before the closest one clips the ray; a callback that collects
    // every hit (returns 1) or skips fixtures (returns -1) needs those and is not served by this call.
    void RayCast(b2RayCastCallback* callback, const b2Vec<float, 2>& point1, const b2Vec<float, 2>& point2);
    int32_t GetBodyCount() const;
    int32_t GetProxyCount() const;
    int32_t GetContactCount() const;
    bool IsLocked() const;
    int32_t GetIndex() const { return m_index; }  // world index within its batch

private:
    friend class b2Body;
    friend class b2Fixture;
    friend class b2WorldBatch;
    b2World(b2WorldBatch* batch, int32_t index, const b2Vec<float, 2>& gravity) : m_batch(batch), m_index(index), m_gravity(gravity) {}
    void SetFlag(int bit, bool on);
    b2WorldBatch* m_batch = nullptr;
    std::unique_ptr<b2WorldBatch> m_owned;  // standalone world: its private batch of one
    int32_t m_index = 0;
    b2Vec<float, 2> m_gravity{{0.0f, -10.0f}};
    int32_t m_flags = B2G_WORLD_DEFAULT_FLAGS;
    std::vector<std::unique_ptr<b2Body>> m_bodies;
    std::vector<std::unique_ptr<b2Fixture>> m_fixtures;
    std::vector<b2g_body_def> m_bodyDefs;
    std::vector<b2g_fixture_def> m_fixtureDefs;
    std::vector<float> m_vertices;
    std::vector<b2g_joint_def> m_jointDefs;
    std::vector<std::unique_ptr<b2Joint>> m_joints;
    std::vector<int32_t> m_firstProxy;  // per fixture: index of its first proxy (creation order)
    int32_t m_proxyCount = 0;
};

// N independent worlds stepped together on one GPU.
class b2WorldBatch {
public:
    b2WorldBatch(const b2Vec<float, 2>& gravity, int32_t worldCount, int32_t device = 0) : m_device(device) {
        if (worldCount <= 0) throw std::runtime_error("b2WorldBatch: worldCount must be positive");
        for (int32_t i = 0; i < worldCount; ++i) m_worlds.emplace_back(new b2World(this, i, gravity));
    }
    ~b2WorldBatch() {
        if (m_handle) b2gpu_batch_destroy(m_handle);
    }
    b2WorldBatch(const b2WorldBatch&) = delete;
    b2WorldBatch& operator=(const b2WorldBatch&) = delete;

    int32_t GetWorldCount() const { return (int32_t)m_worlds.size(); }
    b2World& World(int32_t i) { return *m_worlds[i]; }
    b2World& operator[](int32_t i) { return *m_worlds[i]; }

    // Per-world capacity overrides for the device arenas (0 = derived from the scenes); call before the first Step.
    void SetCapacities(int32_t maxContacts, int32_t maxEvents) { m_maxContacts = maxContacts; m_maxEvents = maxEvents; }
    // Room per world for bodies created after the first Step (b2World::CreateBody on a running world); before the first Step.
    void SetSpawnHeadroom(int32_t bodies, int32_t fixtures, int32_t proxies, int32_t shapeWords, int32_t chainVertices) {
        m_headroom[0] = bodies; m_headroom[1] = fixtures; m_headroom[2] = proxies; m_headroom[3] = shapeWords; m_headroom[4] = chainVertices;
    }
    // Room per world for joints created after the first Step (b2World::CreateJoint on a running world); before the first Step.
    void SetJointHeadroom(int32_t joints) { m_jointHeadroom = joints; }
    void SetContactListener(b2ContactListener* listener) {
        if (m_handle && listener && !m_eventMask) throw std::runtime_error("SetContactListener must precede the first Step");
        m_listener = listener;
    }

    // b2World::Step for every world of the batch.
    void Step(float timeStep, int32_t velocityIterations, int32_t positionIterations) {
        Build();
        Flush();
        Check(b2gpu_batch_step(m_handle, timeStep, velocityIterations, positionIterations));
        m_stateValid = false;
        if (m_listener) DispatchEvents();
    }
    // Step with the dense per-step I/O of an RL / game-server loop fused in (b2gpu_batch_step_io): forces
    // [world][GetMaxBodies()][3] = fx, fy, torque (ApplyForceToCenter / ApplyTorque; may be null) are applied, every
    // world steps, transforms [world][GetMaxBodies()][4] = p.x, p.y, q.s, q.c (may be null) are read back; the host
    // copies of one world group overlap the kernels of the others (pinned buffers).
    void StepIO(const float* forces, float timeStep, int32_t velocityIterations, int32_t positionIterations, float* transforms) {
        Build();
        Flush();
        Check(b2gpu_batch_step_io(m_handle, forces, timeStep, velocityIterations, positionIterations, transforms));
        m_stateValid = false;
        if (m_listener) DispatchEvents();
    }
    int32_t GetMaxBodies() { BuildEarly(); return m_maxBodies; }
    // Which bodies StepIO carries and how (b2gpu_batch_io_select): `forceBodies` / `poseBodies` are creation indices
    // (b2Body::GetIndex(), the same in every world; an empty vector = none, nullptr = every body slot); xya = true reads
    // GetPosition + GetAngle (3 floats per body) instead of GetTransform (4).
    void SelectIO(const std::vector<int32_t>* forceBodies, const std::vector<int32_t>* poseBodies, bool xya) {
        BuildEarly();
        Check(b2gpu_batch_io_select(m_handle, forceBodies ? forceBodies->data() : nullptr, forceBodies ? (int32_t)forceBodies->size() : -1,
                                    poseBodies ? poseBodies->data() : nullptr, poseBodies ? (int32_t)poseBodies->size() : -1,
                                    xya ? B2GPU_POSE_XYA : B2GPU_POSE_XF));
    }
    // `steps` steps replayed from a CUDA graph (no listener dispatch in between).
    void StepN(float timeStep, int32_t velocityIterations, int32_t positionIterations, int32_t steps) {
        Build();
        Flush();
        Check(b2gpu_batch_step_n(m_handle, timeStep, velocityIterations, positionIterations, steps));
        m_stateValid = false;
    }
    void Synchronize() { if (m_handle) Check(b2gpu_batch_sync(m_handle)); }
    bool IsBuilt() const { return m_handle != nullptr; }
    b2gpu_batch* GetHandle() const { return m_handle; }
    // OR of B2GPU_STATUS_* over all worlds (arena overflow is reported, never silent)
    int32_t GetStatus() {
        BuildEarly();
        std::vector<int32_t> st(m_worlds.size());
        Check(b2gpu_batch_download_field(m_handle, B2GPU_F_STATUS, st.data(), 0, (int32_t)st.size()));
        int32_t acc = 0;
        for (int32_t v : st) acc |= v;
        return acc;
    }

private:
    friend class b2World;
    friend class b2Body;
    friend class b2Fixture;
    static void Check(int rc) {
        if (rc != B2GPU_OK) throw std::runtime_error(std::string("b2gpu: ") + b2gpu_last_error());
    }
    // Built before the first Step (an edit or getter that needs the device while the scene was open): the spawn arenas get
    // default headroom so that CreateBody keeps working on the now running worlds.
    void BuildEarly() {
        if (m_handle) return;
        if (m_headroom[0] == 0 && m_headroom[1] == 0 && m_headroom[2] == 0) {
            int32_t nb = 0, nf = 0, np = 0, nv = 0;
            for (auto& w : m_worlds) {
                nb = std::max(nb, (int32_t)w->m_bodyDefs.size());
                nf = std::max(nf, (int32_t)w->m_fixtureDefs.size());
                np = std::max(np, w->m_proxyCount);
                nv = std::max(nv, (int32_t)(w->m_vertices.size() / 2));
            }
            m_headroom[0] = nb / 2 + 16; m_headroom[1] = nf / 2 + 16; m_headroom[2] = np / 2 + 16;
            m_headroom[3] = 64 * (nf / 2 + 16); m_headroom[4] = nv / 2 + 64;
        }
        if (m_jointHeadroom == 0) m_jointHeadroom = 8;
        Build();
    }
    void Build() {
        if (m_handle) return;
        std::vector<b2g_scene> scenes(m_worlds.size());
        for (size_t i = 0; i < m_worlds.size(); ++i) {
            b2World& w = *m_worlds[i];
            b2g_scene& s = scenes[i];
            s.gravity[0] = w.m_gravity[0];
            s.gravity[1] = w.m_gravity[1];
            s.worldFlags = w.m_flags;
            s.bodyCount = (int32_t)w.m_bodyDefs.size();
            s.fixtureCount = (int32_t)w.m_fixtureDefs.size();
            s.vertexCount = (int32_t)(w.m_vertices.size() / 2);
            s.bodies = w.m_bodyDefs.data();
            s.fixtures = w.m_fixtureDefs.data();
            s.vertices = w.m_vertices.data();
            s.joints = w.m_jointDefs.data();
            s.jointCount = (int32_t)w.m_jointDefs.size();
        }
        b2gpu_batch_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        desc.device = m_device;
        desc.numWorlds = (int32_t)m_worlds.size();
        desc.maxContacts = m_maxContacts;
        desc.maxEvents = m_maxEvents;
        m_eventMask = m_listener ? 15 : 0;
        desc.eventMask = m_eventMask;
        Check(b2gpu_scene_capacities(scenes.data(), desc.numWorlds, &desc));
        desc.maxBodies += m_headroom[0]; desc.maxFixtures += m_headroom[1]; desc.maxProxies += m_headroom[2];
        desc.maxShapeWords += m_headroom[3]; desc.maxChainVerts += m_headroom[4];
        desc.maxJoints += m_jointHeadroom;
        if (m_headroom[2] > 0 && m_maxContacts == 0) desc.maxContacts += 4 * m_headroom[2];
        if (desc.eventMask && m_maxEvents == 0 && desc.maxEvents < 3 * desc.maxContacts) desc.maxEvents = 3 * desc.maxContacts;
        Check(b2gpu_batch_create(&desc, &m_handle));
        int rc = b2gpu_batch_build_worlds(m_handle, scenes.data(), desc.numWorlds, 0, 1);
        if (rc != B2GPU_OK) {
            std::string msg = b2gpu_last_error();
            b2gpu_batch_destroy(m_handle);
            m_handle = nullptr;
            throw std::runtime_error("b2gpu: " + msg);
        }
        m_maxBodies = desc.maxBodies;
        m_maxEventsBuilt = desc.maxEvents;
    }
    void Queue(const b2Body* b, int32_t op, bool wake, float vx, float vy, float px, float py);
    void QueueEdit(const b2Body* b, int32_t op, int32_t arg, float vx, float vy, float px, float py);
    void Flush() {  // queued body edits and bodies created on the running worlds -> device, in call order
        FlushCommands();
        FlushSpawns();
    }
    void FlushCommands() {
        if (m_commands.empty()) return;
        Check(b2gpu_batch_body_commands(m_handle, m_commands.data(), (int64_t)m_commands.size()));
        m_commands.clear();
        m_stateValid = false;
    }
    void FlushSpawns() {  // b2World::CreateBody / b2Body::CreateFixture calls made after the first Step
        if (m_spawns.empty()) return;
        std::vector<int32_t> idx(m_spawns.size());
        Check(b2gpu_batch_create_bodies(m_handle, m_spawns.data(), (int32_t)m_spawns.size(), m_spawnFixtures.data(),
                                        m_spawnVertices.data(), idx.data()));
        for (size_t i = 0; i < m_spawns.size(); ++i)
            if (idx[i] != m_spawnBodies[i]->GetIndex()) throw std::runtime_error("b2gpu: a world's arenas are full (SetSpawnHeadroom)");
        m_spawns.clear(); m_spawnFixtures.clear(); m_spawnVertices.clear(); m_spawnBodies.clear();
        m_stateValid = false;
    }
    void Refresh() {  // lazy readback of body state after a Step
        Build();
        Flush();
        if (m_stateValid) return;
        const size_t n = m_worlds.size() * (size_t)m_maxBodies;
        m_xf.resize(n * 4); m_sweep.resize(n * 4); m_vel.resize(n * 4); m_bodyFlags.resize(n);
        const int32_t W = (int32_t)m_worlds.size();
        Check(b2gpu_batch_download_field(m_handle, B2GPU_F_BODY_XF, m_xf.data(), 0, W));
        Check(b2gpu_batch_download_field(m_handle, B2GPU_F_BODY_SWEEP, m_sweep.data(), 0, W));
        Check(b2gpu_batch_download_field(m_handle, B2GPU_F_BODY_VEL, m_vel.data(), 0, W));
        Check(b2gpu_batch_download_field(m_handle, B2GPU_F_BODY_FLAGS, m_bodyFlags.data(), 0, W));
        m_stateValid = true;
    }
    size_t Slot(const b2Body* b) const { return (size_t)b->m_world->m_index * m_maxBodies + b->m_index; }
    void DispatchEvents() {  // compacted readback: only the events that happened cross the bus
        const int32_t W = (int32_t)m_worlds.size();
        m_eventOffsets.resize((size_t)W + 1);
        Check(b2gpu_batch_read_events(m_handle, m_eventOffsets.data(), nullptr, 0));
        m_events.resize((size_t)m_eventOffsets[W]);
        if (m_events.empty()) return;
        Check(b2gpu_batch_read_events(m_handle, m_eventOffsets.data(), m_events.data(), (int64_t)m_events.size()));
        for (int32_t w = 0; w < W; ++w) {
            b2World& world = *m_worlds[w];
            for (int64_t i = m_eventOffsets[w]; i < m_eventOffsets[w + 1]; ++i) {
                const b2gpu_event& e = m_events[(size_t)i];
                b2ContactEvent ev{&world, world.m_fixtures[e.fixtureA].get(), world.m_fixtures[e.fixtureB].get(), e.childA, e.childB,
                                  e.count};
                switch (e.type) {
                    case B2GPU_EVENT_BEGIN: m_listener->BeginContact(ev); break;
                    case B2GPU_EVENT_END: m_listener->EndContact(ev); break;
                    case B2GPU_EVENT_PRESOLVE: m_listener->PreSolve(ev); break;
                    default: {
                        b2ContactImpulse imp{{e.normalImpulses[0], e.normalImpulses[1]}, {e.tangentImpulses[0], e.tangentImpulses[1]},
                                             e.count};
                        m_listener->PostSolve(ev, imp);
                    }
                }
            }
        }
    }

    int32_t m_device = 0;
    std::vector<std::unique_ptr<b2World>> m_worlds;
    b2gpu_batch* m_handle = nullptr;
    b2ContactListener* m_listener = nullptr;
    int32_t m_eventMask = 0, m_maxContacts = 0, m_maxEvents = 0, m_maxBodies = 0, m_maxEventsBuilt = 0;
    bool m_stateValid = false;
    std::vector<float> m_xf, m_sweep, m_vel;
    std::vector<b2gpu_body_command> m_commands;
    int32_t m_headroom[5] = {0, 0, 0, 0, 0};
    int32_t m_jointHeadroom = 0;
    bool m_destroyedWhileOpen = false;  // DestroyBody before the first Step (see b2World::CreateBody)
    // bodies created after the first Step, not yet on the device (only commands OR spawns are pending at any time, so
    // that the device sees creations and edits in call order)
    std::vector<b2gpu_spawn> m_spawns;
    std::vector<b2g_fixture_def> m_spawnFixtures;
    std::vector<float> m_spawnVertices;
    std::vector<b2Body*> m_spawnBodies;
    std::vector<int64_t> m_eventOffsets;
    std::vector<b2gpu_event> m_events;
    std::vector<int32_t> m_bodyFlags;
};

// ---- inline members ---------------------------------------------------------------------------------
inline b2World::b2World(const b2Vec<float, 2>& gravity) : m_gravity(gravity) {
    m_owned.reset(new b2WorldBatch(gravity, 1));
    m_batch = m_owned.get();
}
inline b2World::~b2World() {}

// A standalone b2World forwards to world 0 of its private batch, which holds the recorded scene.
#define B2_SELF (m_owned ? m_owned->World(0) : *this)

inline void b2World::SetFlag(int bit, bool on) {
    b2World& w = B2_SELF;
    if (w.m_batch->IsBuilt()) throw std::runtime_error("world flags are frozen at the first Step");
    w.m_flags = on ? (w.m_flags | bit) : (w.m_flags & ~bit);
    if (&w != this) m_flags = w.m_flags;
}
inline bool b2World::IsLocked() const { return m_batch->IsBuilt(); }
inline int32_t b2World::GetBodyCount() const { return (int32_t)(m_owned ? m_owned->World(0) : *this).m_bodies.size(); }
inline b2Body* b2World::CreateBody(const b2BodyDef* def) {
    b2World& w = B2_SELF;
    // A body was destroyed while the scene was still open: what is created after it must see the freed proxy ids
    // (b2DynamicTree's LIFO free list), so the batch is built now and this body takes the running-world path.
    if (!w.m_batch->IsBuilt() && w.m_batch->m_destroyedWhileOpen) w.m_batch->BuildEarly();
    const bool running = w.m_batch->IsBuilt();  // the world has been stepped: the body goes to the device at the next flush
    b2g_body_def d;
    std::memset(&d, 0, sizeof(d));
    d.type = (int32_t)def->type;
    d.flags = (def->allowSleep ? B2G_BD_ALLOW_SLEEP : 0) | (def->awake ? B2G_BD_AWAKE : 0) |
              (def->fixedRotation ? B2G_BD_FIXED_ROTATION : 0) | (def->bullet ? B2G_BD_BULLET : 0) |
              (def->active ? B2G_BD_ACTIVE : 0);
    d.position[0] = def->position[0]; d.position[1] = def->position[1];
    d.angle = def->angle;
    d.linearVelocity[0] = def->linearVelocity[0]; d.linearVelocity[1] = def->linearVelocity[1];
    d.angularVelocity = def->angularVelocity;
    d.linearDamping = def->linearDamping;
    d.angularDamping = def->angularDamping;
    d.gravityScale = def->gravityScale;
    w.m_bodyDefs.push_back(d);
    std::unique_ptr<b2Body> b(new b2Body);
    b->m_world = &w;
    b->m_index = (int32_t)w.m_bodies.size();
    b->m_type = def->type;
    b->m_userData = def->userData;
    b->m_linearDamping = def->linearDamping;
    b->m_angularDamping = def->angularDamping;
    b->m_shadow.p[0] = d.position[0]; b->m_shadow.p[1] = d.position[1]; b->m_shadow.angle = d.angle;
    b->m_shadow.v[0] = d.linearVelocity[0]; b->m_shadow.v[1] = d.linearVelocity[1]; b->m_shadow.w = d.angularVelocity;
    b->m_shadow.awake = def->awake;
    w.m_bodies.push_back(std::move(b));
    if (running) {
        b2WorldBatch* B = w.m_batch;
        B->FlushCommands();  // edits made before this creation come first
        b2gpu_spawn sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.world = w.m_index;
        sp.firstFixture = (int32_t)B->m_spawnFixtures.size();
        sp.body = d;
        B->m_spawns.push_back(sp);
        B->m_spawnBodies.push_back(w.m_bodies.back().get());
    }
    return w.m_bodies.back().get();
}
inline b2Joint* b2World::CreateJoint(const b2JointDef* def) {
    b2World& w = B2_SELF;
    if (!def->bodyA || !def->bodyB) return nullptr;
    const bool running = w.m_batch->IsBuilt();
    b2g_joint_def d;
    std::memset(&d, 0, sizeof(d));
    d.bodyA = def->bodyA->GetIndex();
    d.bodyB = def->bodyB->GetIndex();
    d.collideConnected = def->collideConnected ? 1 : 0;
    auto set2 = [](float* dst, const b2Vec<float, 2>& v) { dst[0] = v[0]; dst[1] = v[1]; };
    if (def->type == b2JointType::REVOLUTE_JOINT) {
        const b2RevoluteJointDef* r = static_cast<const b2RevoluteJointDef*>(def);
        d.type = B2G_JOINT_REVOLUTE;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB);
        d.referenceAngle = r->referenceAngle;
        d.enableLimit = r->enableLimit; d.lower = r->lowerAngle; d.upper = r->upperAngle;
        d.enableMotor = r->enableMotor; d.maxMotor = r->maxMotorTorque; d.motorSpeed = r->motorSpeed;
    } else if (def->type == b2JointType::PRISMATIC_JOINT) {
        const b2PrismaticJointDef* r = static_cast<const b2PrismaticJointDef*>(def);
        d.type = B2G_JOINT_PRISMATIC;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB); set2(d.localAxisA, r->localAxisA);
        d.referenceAngle = r->referenceAngle;
        d.enableLimit = r->enableLimit; d.lower = r->lowerTranslation; d.upper = r->upperTranslation;
        d.enableMotor = r->enableMotor; d.maxMotor = r->maxMotorForce; d.motorSpeed = r->motorSpeed;
    } else if (def->type == b2JointType::DISTANCE_JOINT) {
        const b2DistanceJointDef* r = static_cast<const b2DistanceJointDef*>(def);
        d.type = B2G_JOINT_DISTANCE;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB);
        d.length = r->length; d.frequencyHz = r->frequencyHz; d.dampingRatio = r->dampingRatio;
    } else if (def->type == b2JointType::WHEEL_JOINT) {
        const b2WheelJointDef* r = static_cast<const b2WheelJointDef*>(def);
        d.type = B2G_JOINT_WHEEL;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB); set2(d.localAxisA, r->localAxisA);
        d.enableMotor = r->enableMotor; d.maxMotor = r->maxMotorTorque; d.motorSpeed = r->motorSpeed;
        d.frequencyHz = r->frequencyHz; d.dampingRatio = r->dampingRatio;
    } else if (def->type == b2JointType::PULLEY_JOINT) {
        const b2PulleyJointDef* r = static_cast<const b2PulleyJointDef*>(def);
        d.type = B2G_JOINT_PULLEY;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB); set2(d.localAxisA, r->groundAnchorA);
        d.lower = r->groundAnchorB[0]; d.upper = r->groundAnchorB[1];
        d.length = r->lengthA; d.maxMotor = r->lengthB; d.motorSpeed = r->ratio;
    } else if (def->type == b2JointType::MOUSE_JOINT) {
        const b2MouseJointDef* r = static_cast<const b2MouseJointDef*>(def);
        d.type = B2G_JOINT_MOUSE;
        set2(d.localAnchorA, r->target);
        d.length = r->maxForce; d.frequencyHz = r->frequencyHz; d.dampingRatio = r->dampingRatio;
    } else if (def->type == b2JointType::GEAR_JOINT) {
        const b2GearJointDef* r = static_cast<const b2GearJointDef*>(def);
        if (!r->joint1 || !r->joint2) return nullptr;
        d.type = B2G_JOINT_GEAR;
        d.enableLimit = r->joint1->GetIndex(); d.enableMotor = r->joint2->GetIndex();
        d.motorSpeed = r->ratio;
        // b2GearJoint.cpp:59-60, 92-93: the gear's bodies are the second bodies of joint1 / joint2
        d.bodyA = r->joint1->GetBodyB()->GetIndex(); d.bodyB = r->joint2->GetBodyB()->GetIndex();
    } else if (def->type == b2JointType::MOTOR_JOINT) {
        const b2MotorJointDef* r = static_cast<const b2MotorJointDef*>(def);
        d.type = B2G_JOINT_MOTOR;
        set2(d.localAnchorA, r->linearOffset);
        d.referenceAngle = r->angularOffset;
        d.length = r->maxForce; d.maxMotor = r->maxTorque; d.frequencyHz = r->correctionFactor;
    } else if (def->type == b2JointType::ROPE_JOINT) {
        const b2RopeJointDef* r = static_cast<const b2RopeJointDef*>(def);
        d.type = B2G_JOINT_ROPE;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB);
        d.length = r->maxLength;
    } else if (def->type == b2JointType::FRICTION_JOINT) {
        const b2FrictionJointDef* r = static_cast<const b2FrictionJointDef*>(def);
        d.type = B2G_JOINT_FRICTION;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB);
        d.length = r->maxForce; d.maxMotor = r->maxTorque;
    } else if (def->type == b2JointType::WELD_JOINT) {
        const b2WeldJointDef* r = static_cast<const b2WeldJointDef*>(def);
        d.type = B2G_JOINT_WELD;
        set2(d.localAnchorA, r->localAnchorA); set2(d.localAnchorB, r->localAnchorB);
        d.referenceAngle = r->referenceAngle;
        d.frequencyHz = r->frequencyHz; d.dampingRatio = r->dampingRatio;
    } else {
        return nullptr;
    }
    if (running) {  // b2World::CreateJoint on the running world, after everything queued before it
        b2WorldBatch* B = w.m_batch;
        B->Flush();
        int32_t id = -1;
        b2WorldBatch::Check(b2gpu_batch_create_joints(B->m_handle, &w.m_index, &d, 1, &id));
        if (id != (int32_t)w.m_joints.size()) throw std::runtime_error("b2gpu: the world's joint arena is full (SetJointHeadroom)");
        B->m_stateValid = false;
    } else {
        w.m_jointDefs.push_back(d);
    }
    std::unique_ptr<b2Joint> j(new b2Joint);
    j->m_type = def->type;
    j->m_bodyA = w.m_bodies[d.bodyA].get();
    j->m_bodyB = w.m_bodies[d.bodyB].get();
    j->m_userData = def->userData;
    j->m_index = (int32_t)w.m_joints.size();
    w.m_joints.push_back(std::move(j));
    return w.m_joints.back().get();
}
inline void b2World::DestroyJoint(b2Joint* joint) {
    b2World& w = B2_SELF;
    if (!joint || joint->GetIndex() < 0 || joint->GetIndex() >= (int32_t)w.m_joints.size() || w.m_joints[joint->GetIndex()].get() != joint) return;
    if (!joint->m_bodyA && !joint->m_bodyB) return;  // destroyed already
    b2WorldBatch* B = w.m_batch;
    B->BuildEarly();
    B->FlushSpawns();
    b2gpu_body_command c;
    std::memset(&c, 0, sizeof(c));
    c.world = w.m_index; c.op = B2GPU_CMD_DESTROY_JOINT; c.wake = joint->GetIndex();
    B->m_commands.push_back(c);
    joint->m_bodyA = joint->m_bodyB = nullptr;  // destroyed: a second DestroyJoint / SetTarget on it does nothing
}
inline void b2World::DestroyBody(b2Body* body) {
    if (!body || body->GetWorld() != &(B2_SELF)) return;
    m_batch->Queue(body, B2GPU_CMD_DESTROY_BODY, false, 0.0f, 0.0f, 0.0f, 0.0f);
}
inline void b2World::Step(float dt, int32_t vi, int32_t pi) { m_batch->Step(dt, vi, pi); }
inline void b2World::SetContactListener(b2ContactListener* l) { m_batch->SetContactListener(l); }
inline void b2World::QueryAABB(b2QueryCallback* callback, const b2AABB& aabb) {
    b2World& w = B2_SELF;
    b2WorldBatch* B = w.m_batch;
    B->BuildEarly();
    B->Flush();
    b2gpu_aabb_query q;
    q.world = w.m_index;
    q.lower[0] = aabb.lowerBound[0]; q.lower[1] = aabb.lowerBound[1];
    q.upper[0] = aabb.upperBound[0]; q.upper[1] = aabb.upperBound[1];
    std::vector<int32_t> hits((size_t)std::max(w.m_proxyCount, 1));
    int32_t count = 0;
    b2WorldBatch::Check(b2gpu_batch_query_aabb(B->m_handle, &q, 1, (int32_t)hits.size(), &count, hits.data()));
    count = std::min(count, (int32_t)hits.size());
    for (int32_t i = 0; i < count; ++i) {
        const int32_t fixture = hits[(size_t)i] & 0xffff;
        if (fixture >= (int32_t)w.m_fixtures.size()) continue;
        if (!callback->ReportFixture(w.m_fixtures[(size_t)fixture].get())) break;
    }
}
inline void b2World::RayCast(b2RayCastCallback* callback, const b2Vec<float, 2>& p1, const b2Vec<float, 2>& p2) {
    b2World& w = B2_SELF;
    b2WorldBatch* B = w.m_batch;
    B->BuildEarly();
    B->Flush();
    b2gpu_ray ray;
    ray.world = w.m_index;
    ray.p1[0] = p1[0]; ray.p1[1] = p1[1]; ray.p2[0] = p2[0]; ray.p2[1] = p2[1];
    ray.maxFraction = 1.0f;
    b2gpu_ray_hit hit;
    b2WorldBatch::Check(b2gpu_batch_raycast_closest(B->m_handle, &ray, 1, &hit));
    if (hit.fixture < 0 || hit.fixture >= (int32_t)w.m_fixtures.size()) return;
    callback->ReportFixture(w.m_fixtures[(size_t)hit.fixture].get(), b2Vec<float, 2>{{hit.point[0], hit.point[1]}},
                            b2Vec<float, 2>{{hit.normal[0], hit.normal[1]}}, hit.fraction);
}
inline int32_t b2World::GetProxyCount() const { return (m_owned ? m_owned->World(0) : *this).m_proxyCount; }
inline int32_t b2World::GetContactCount() const {
    b2WorldBatch* batch = m_batch;
    const b2World& w = m_owned ? m_owned->World(0) : *this;
    if (!batch->IsBuilt()) return 0;  // contacts are created by the first Step (b2World::Step -> FindNewContacts)
    int32_t n = 0;
    b2WorldBatch::Check(b2gpu_batch_download_field(batch->m_handle, B2GPU_F_CONTACT_COUNT, &n, w.m_index, 1));
    return n;
}
#undef B2_SELF

inline b2Fixture* b2Body::CreateFixture(const b2FixtureDef* def) {
    b2World& w = *m_world;
    if (!def->shape) return nullptr;
    b2WorldBatch* B = w.m_batch;
    const bool running = B->IsBuilt();
    // on a running world the body created last (nothing else queued since) takes its fixtures in its spawn record
    // (b2gpu_batch_create_bodies); any other body gets them through b2gpu_batch_create_fixtures, at once
    const bool attach = running && (B->m_spawnBodies.empty() || B->m_spawnBodies.back() != this);
    b2g_fixture_def fd;
    std::memset(&fd, 0, sizeof(fd));
    fd.body = m_index;
    fd.isSensor = def->isSensor;
    fd.friction = def->friction;
    fd.restitution = def->restitution;
    fd.density = def->density;
    fd.categoryBits = def->filter.categoryBits;
    fd.maskBits = def->filter.maskBits;
    fd.groupIndex = def->filter.groupIndex;
    if (attach) {
        B->Flush();
        std::vector<float> verts;
        def->shape->Describe(fd, verts);
        b2gpu_attach at;
        at.world = w.m_index; at.body = m_index; at.firstFixture = 0; at.fixtureCount = 1;
        int32_t idx = -1;
        b2WorldBatch::Check(b2gpu_batch_create_fixtures(B->m_handle, &at, 1, &fd, verts.empty() ? nullptr : verts.data(), &idx));
        if (idx != (int32_t)w.m_fixtures.size()) throw std::runtime_error("b2gpu: a world's arenas are full (SetSpawnHeadroom)");
        B->m_stateValid = false;
    } else if (running) {
        def->shape->Describe(fd, B->m_spawnVertices);
        B->m_spawnFixtures.push_back(fd);
        B->m_spawns.back().fixtureCount += 1;
    } else {
        def->shape->Describe(fd, w.m_vertices);
        w.m_fixtureDefs.push_back(fd);
    }
    std::unique_ptr<b2Fixture> f(new b2Fixture);
    f->m_body = this;
    f->m_userData = def->userData;
    f->m_friction = def->friction;
    f->m_restitution = def->restitution;
    f->m_type = def->shape->GetType();
    f->m_def = fd;
    def->shape->Describe(f->m_def, f->m_defVertices);  // its own copy of the shape: SetDensity recomputes the mass data from it
    f->m_index = (int32_t)w.m_fixtures.size();
    w.m_firstProxy.push_back(w.m_proxyCount);
    // proxies: one per shape child (b2Fixture::CreateProxies, b2Fixture.cpp:124-139)
    int children = 1;
    if (fd.shapeType == B2G_SHAPE_CHAIN) children = fd.shapeKind == B2G_CHAIN_LOOP ? fd.vertexCount : fd.vertexCount - 1;
    w.m_proxyCount += children;
    m_shadow.fixtures += 1;
    if (def->density > 0.0f && m_shadow.w != 0.0f) m_shadow.exact = false;  // ResetMassData moves the velocity (b2Body.cpp:350)
    w.m_fixtures.push_back(std::move(f));
    return w.m_fixtures.back().get();
}
inline const b2Vec<float, 2>& b2Body::GetPosition() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) {  // scene still open (joint defs' Initialize())
        m_tmp[0] = {{m_shadow.p[0], m_shadow.p[1]}};
        return m_tmp[0];
    }
    B->Refresh();
    const float* x = &B->m_xf[B->Slot(this) * 4];
    m_tmp[0] = {{x[0], x[1]}};
    return m_tmp[0];
}
inline float b2Body::GetAngle() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) return m_shadow.angle;  // scene still open (joint defs' Initialize())
    B->Refresh();
    return B->m_sweep[B->Slot(this) * 4 + 2];
}
inline const b2Vec<float, 2>& b2Body::GetWorldCenter() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) {  // scene still open: position + R(angle) * localCenter, the centre from the fixtures recorded so far
        b2World& w = *m_world;
        if (m_shadow.massFixtures != (int32_t)w.m_fixtureDefs.size()) {
            b2g_scene sc;
            std::memset(&sc, 0, sizeof(sc));
            sc.bodyCount = (int32_t)w.m_bodyDefs.size(); sc.bodies = w.m_bodyDefs.data();
            sc.fixtureCount = (int32_t)w.m_fixtureDefs.size(); sc.fixtures = w.m_fixtureDefs.data();
            sc.vertexCount = (int32_t)(w.m_vertices.size() / 2); sc.vertices = w.m_vertices.data();
            float md[4];
            b2WorldBatch::Check(b2gpu_host_body_mass_data(&sc, m_index, md));
            m_shadow.localCenter[0] = md[2]; m_shadow.localCenter[1] = md[3];
            m_shadow.massFixtures = sc.fixtureCount;
        }
        float sn, cs;
        b2WorldBatch::Check(b2gpu_host_sincosf(&m_shadow.angle, &sn, &cs, 1));  // b2Rot::Set as the library evaluates it
        const float lx = m_shadow.localCenter[0], ly = m_shadow.localCenter[1];
        m_tmp[1] = {{(cs * lx - sn * ly) + m_shadow.p[0], (sn * lx + cs * ly) + m_shadow.p[1]}};
        return m_tmp[1];
    }
    B->Refresh();
    const float* s = &B->m_sweep[B->Slot(this) * 4];
    m_tmp[1] = {{s[0], s[1]}};
    return m_tmp[1];
}
inline const b2Vec<float, 2>& b2Body::GetLinearVelocity() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt() && m_shadow.exact) {
        m_tmp[2] = {{m_shadow.v[0], m_shadow.v[1]}};
        return m_tmp[2];
    }
    if (!B->IsBuilt()) B->BuildEarly();
    B->Refresh();
    const float* v = &B->m_vel[B->Slot(this) * 4];
    m_tmp[2] = {{v[0], v[1]}};
    return m_tmp[2];
}
inline float b2Body::GetAngularVelocity() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt() && m_shadow.exact) return m_shadow.w;
    if (!B->IsBuilt()) B->BuildEarly();
    B->Refresh();
    return B->m_vel[B->Slot(this) * 4 + 2];
}
inline float b2Body::GetMass() const {
    b2WorldBatch* B = m_world->m_batch;
    b2World& w = *m_world;
    if (!B->IsBuilt()) {  // scene still open: b2Body::ResetMassData over the fixtures recorded so far, on the host
        b2g_scene sc;
        std::memset(&sc, 0, sizeof(sc));
        sc.bodyCount = (int32_t)w.m_bodyDefs.size(); sc.bodies = w.m_bodyDefs.data();
        sc.fixtureCount = (int32_t)w.m_fixtureDefs.size(); sc.fixtures = w.m_fixtureDefs.data();
        sc.vertexCount = (int32_t)(w.m_vertices.size() / 2); sc.vertices = w.m_vertices.data();
        float md[4];
        b2WorldBatch::Check(b2gpu_host_body_mass_data(&sc, m_index, md));
        return md[0];
    }
    B->Flush();
    std::vector<float> damp((size_t)B->m_maxBodies * 4);  // this world's B2GPU_F_BODY_DAMP rows: .w = m_mass
    b2WorldBatch::Check(b2gpu_batch_download_field(B->m_handle, B2GPU_F_BODY_DAMP, damp.data(), w.m_index, 1));
    return damp[(size_t)m_index * 4 + 3];
}
inline bool b2Body::IsAwake() const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) return m_shadow.awake;
    B->Refresh();
    return (B->m_bodyFlags[B->Slot(this)] & 8) != 0;
}
// Transform of a body: from its definition until the batch is built, from the device afterwards.
inline b2Vec<float, 2> b2Body::GetLocalPoint(const b2Vec<float, 2>& p) const {
    float px, py, qs, qc;
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) {
        px = m_shadow.p[0]; py = m_shadow.p[1]; qs = std::sin(m_shadow.angle); qc = std::cos(m_shadow.angle);
    } else {
        B->Refresh();
        const float* x = &B->m_xf[B->Slot(this) * 4];
        px = x[0]; py = x[1]; qs = x[2]; qc = x[3];
    }
    const float dx = p[0] - px, dy = p[1] - py;  // b2MulT(xf, p)
    return b2Vec<float, 2>{{qc * dx + qs * dy, -qs * dx + qc * dy}};
}
inline b2Vec<float, 2> b2Body::GetLocalVector(const b2Vec<float, 2>& v) const {
    float qs, qc;
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) {
        qs = std::sin(m_shadow.angle); qc = std::cos(m_shadow.angle);
    } else {
        B->Refresh();
        const float* x = &B->m_xf[B->Slot(this) * 4];
        qs = x[2]; qc = x[3];
    }
    return b2Vec<float, 2>{{qc * v[0] + qs * v[1], -qs * v[0] + qc * v[1]}};  // b2MulT(q, v)
}
inline b2Vec<float, 2> b2Body::GetWorldPoint(const b2Vec<float, 2>& p) const {
    const b2Vec<float, 2> x = GetLocalVector(b2Vec<float, 2>{{1.0f, 0.0f}});  // (q.c, -q.s)
    const b2Vec<float, 2>& o = GetPosition();
    const float qc = x[0], qs = -x[1];
    return b2Vec<float, 2>{{(qc * p[0] - qs * p[1]) + o[0], (qs * p[0] + qc * p[1]) + o[1]}};  // b2Mul(xf, p)
}
inline b2Vec<float, 2> b2Body::GetWorldVector(const b2Vec<float, 2>& v) const {
    const b2Vec<float, 2> x = GetLocalVector(b2Vec<float, 2>{{1.0f, 0.0f}});
    const float qc = x[0], qs = -x[1];
    return b2Vec<float, 2>{{qc * v[0] - qs * v[1], qs * v[0] + qc * v[1]}};  // b2Mul(q, v)
}
inline b2Vec<float, 2> b2Body::GetLinearVelocityFromWorldPoint(const b2Vec<float, 2>& wp) const {
    const b2Vec<float, 2> c = GetWorldCenter(), v = GetLinearVelocity();
    const float w = GetAngularVelocity();
    const float dx = wp[0] - c[0], dy = wp[1] - c[1];
    return b2Vec<float, 2>{{v[0] + (-w * dy), v[1] + w * dx}};  // m_linearVelocity + b2Cross(m_angularVelocity, worldPoint - m_sweep.c)
}
inline bool b2Body::HasFlag(int deviceBit, int defBit) const {
    b2WorldBatch* B = m_world->m_batch;
    if (!B->IsBuilt()) return (m_world->m_bodyDefs[(size_t)m_index].flags & defBit) != 0;
    B->Refresh();
    return (B->m_bodyFlags[B->Slot(this)] & deviceBit) != 0;
}
// The defs' Initialize() helpers (b2RevoluteJoint.cpp:38-45, b2PrismaticJoint.cpp:90-98, b2DistanceJoint.cpp:40-49)
inline void b2RevoluteJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& anchor) {
    bodyA = bA; bodyB = bB;
    localAnchorA = bodyA->GetLocalPoint(anchor);
    localAnchorB = bodyB->GetLocalPoint(anchor);
    referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}
inline void b2PrismaticJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& anchor, const b2Vec<float, 2>& axis) {
    bodyA = bA; bodyB = bB;
    localAnchorA = bodyA->GetLocalPoint(anchor);
    localAnchorB = bodyB->GetLocalPoint(anchor);
    localAxisA = bodyA->GetLocalVector(axis);
    referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}
inline void b2WheelJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& anchor, const b2Vec<float, 2>& axis) {
    bodyA = bA; bodyB = bB;
    localAnchorA = bodyA->GetLocalPoint(anchor);
    localAnchorB = bodyB->GetLocalPoint(anchor);
    localAxisA = bodyA->GetLocalVector(axis);
}
inline void b2PulleyJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& groundA, const b2Vec<float, 2>& groundB,
                                         const b2Vec<float, 2>& anchorA, const b2Vec<float, 2>& anchorB, float r) {
    bodyA = bA; bodyB = bB;
    groundAnchorA = groundA; groundAnchorB = groundB;
    localAnchorA = bodyA->GetLocalPoint(anchorA);
    localAnchorB = bodyB->GetLocalPoint(anchorB);
    const float ax = anchorA[0] - groundA[0], ay = anchorA[1] - groundA[1];
    lengthA = std::sqrt(ax * ax + ay * ay);
    const float bx = anchorB[0] - groundB[0], by = anchorB[1] - groundB[1];
    lengthB = std::sqrt(bx * bx + by * by);
    ratio = r;
}
inline void b2MotorJointDef::Initialize(b2Body* bA, b2Body* bB) {
    bodyA = bA; bodyB = bB;
    linearOffset = bodyA->GetLocalPoint(bodyB->GetPosition());
    angularOffset = bodyB->GetAngle() - bodyA->GetAngle();
}
inline void b2FrictionJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& anchor) {
    bodyA = bA; bodyB = bB;
    localAnchorA = bodyA->GetLocalPoint(anchor);
    localAnchorB = bodyB->GetLocalPoint(anchor);
}
inline void b2WeldJointDef::Initialize(b2Body* bA, b2Body* bB, const b2Vec<float, 2>& anchor) {
    bodyA = bA; bodyB = bB;
    localAnchorA = bodyA->GetLocalPoint(anchor);
    localAnchorB = bodyB->GetLocalPoint(anchor);
    referenceAngle = bodyB->GetAngle() - bodyA->GetAngle();
}
inline void b2DistanceJointDef::Initialize(b2Body* b1, b2Body* b2, const b2Vec<float, 2>& anchor1, const b2Vec<float, 2>& anchor2) {
    bodyA = b1; bodyB = b2;
    localAnchorA = bodyA->GetLocalPoint(anchor1);
    localAnchorB = bodyB->GetLocalPoint(anchor2);
    const float dx = anchor2[0] - anchor1[0], dy = anchor2[1] - anchor1[1];
    length = std::sqrt(dx * dx + dy * dy);
}

inline void b2WorldBatch::Queue(const b2Body* b, int32_t op, bool wake, float vx, float vy, float px, float py) {
    if (m_handle) FlushSpawns();  // a body created earlier must exist before an edit that follows it
    b2gpu_body_command c;
    c.world = b->m_world->m_index;
    c.body = b->m_index;
    c.op = op;
    c.wake = wake ? 1 : 0;
    c.v[0] = vx; c.v[1] = vy; c.p[0] = px; c.p[1] = py;
    if (m_handle) {
        m_commands.push_back(c);
        return;
    }
    // Scene still open.  A body without fixtures has no proxies and no mass data yet, so a velocity / transform / awake
    // setter is exactly a change of its definition (and b2Body::ResetMassData of the fixtures that follow adjusts the
    // velocity as the reference does, b2Body.cpp:350).  Everything else is queued and applied by the device, in call
    // order, right after the batch is built.  (Not reproduced: a fixture added to a body AFTER a setter that itself
    // followed an earlier fixture of that body.)
    b2g_body_def& def = b->m_world->m_bodyDefs[b->m_index];
    const bool bare = b->m_shadow.fixtures == 0;
    bool folded = false;
    if (bare && b->m_type != b2BodyType::STATIC_BODY && op == B2GPU_CMD_SET_LINEAR_VELOCITY) {
        def.linearVelocity[0] = vx; def.linearVelocity[1] = vy;
        if (vx * vx + vy * vy > 0.0f) def.flags |= B2G_BD_AWAKE;
        folded = true;
    } else if (bare && b->m_type != b2BodyType::STATIC_BODY && op == B2GPU_CMD_SET_ANGULAR_VELOCITY) {
        def.angularVelocity = vx;
        if (vx * vx > 0.0f) def.flags |= B2G_BD_AWAKE;
        folded = true;
    } else if (bare && op == B2GPU_CMD_SET_TRANSFORM) {
        def.position[0] = vx; def.position[1] = vy; def.angle = px;
        folded = true;
    } else if (bare && op == B2GPU_CMD_SET_AWAKE) {
        if (wake) def.flags |= B2G_BD_AWAKE;
        else { def.flags &= ~B2G_BD_AWAKE; def.linearVelocity[0] = def.linearVelocity[1] = def.angularVelocity = 0.0f; }
        folded = true;
    } else if ((op == B2GPU_CMD_SET_LINEAR_VELOCITY || op == B2GPU_CMD_SET_ANGULAR_VELOCITY) && b->m_type == b2BodyType::STATIC_BODY) {
        folded = true;  // b2Body::SetLinearVelocity on a static body does nothing
    }
    if (!folded) m_commands.push_back(c);
    // keep the host shadow in step (b2Body.h:408-467, 634-653; b2Body.cpp:418-440)
    b2Body::Shadow& s = b->m_shadow;
    const bool isStatic = b->m_type == b2BodyType::STATIC_BODY;
    switch (op) {
        case B2GPU_CMD_SET_LINEAR_VELOCITY:
            if (isStatic) break;
            if (vx * vx + vy * vy > 0.0f) s.awake = true;
            s.v[0] = vx; s.v[1] = vy;
            break;
        case B2GPU_CMD_SET_ANGULAR_VELOCITY:
            if (isStatic) break;
            if (vx * vx > 0.0f) s.awake = true;
            s.w = vx;
            break;
        case B2GPU_CMD_SET_TRANSFORM:
            s.p[0] = vx; s.p[1] = vy; s.angle = px;
            break;
        case B2GPU_CMD_SET_AWAKE:
            if (wake) s.awake = true;
            else { s.awake = false; s.v[0] = s.v[1] = s.w = 0.0f; }
            break;
        case B2GPU_CMD_DESTROY_BODY:
            s.destroyed = true;
            m_destroyedWhileOpen = true;
            break;
        case B2GPU_CMD_APPLY_LINEAR_IMPULSE:
        case B2GPU_CMD_APPLY_ANGULAR_IMPULSE:
            s.exact = false;  // the new velocity needs the body's mass data: velocity getters build the batch
            break;
        default:  // forces, torques, the mouse target: no getter of the mirror shows them before the step
            if (wake && !isStatic) s.awake = true;
            break;
    }
}
inline void b2Body::SetMouseTarget(const b2Vec<float, 2>& target) {
    m_world->m_batch->Queue(this, B2GPU_CMD_MOUSE_SET_TARGET, true, target[0], target[1], 0.0f, 0.0f);
}
inline void b2Joint::SetTarget(const b2Vec<float, 2>& target) {
    if (m_type == b2JointType::MOUSE_JOINT && m_bodyB) m_bodyB->SetMouseTarget(target);
}
inline void b2Body::ApplyForce(const b2Vec<float, 2>& force, const b2Vec<float, 2>& point, bool wake) {
    m_world->m_batch->Queue(this, B2GPU_CMD_APPLY_FORCE, wake, force[0], force[1], point[0], point[1]);
}
inline void b2Body::ApplyForceToCenter(const b2Vec<float, 2>& force, bool wake) {
    m_world->m_batch->Queue(this, B2GPU_CMD_APPLY_FORCE_TO_CENTER, wake, force[0], force[1], 0.0f, 0.0f);
}
inline void b2Body::ApplyTorque(float torque, bool wake) {
    m_world->m_batch->Queue(this, B2GPU_CMD_APPLY_TORQUE, wake, torque, 0.0f, 0.0f, 0.0f);
}
inline void b2Body::ApplyLinearImpulse(const b2Vec<float, 2>& impulse, const b2Vec<float, 2>& point, bool wake) {
    m_world->m_batch->Queue(this, B2GPU_CMD_APPLY_LINEAR_IMPULSE, wake, impulse[0], impulse[1], point[0], point[1]);
}
inline void b2Body::ApplyAngularImpulse(float impulse, bool wake) {
    m_world->m_batch->Queue(this, B2GPU_CMD_APPLY_ANGULAR_IMPULSE, wake, impulse, 0.0f, 0.0f, 0.0f);
}
inline void b2Body::SetLinearVelocity(const b2Vec<float, 2>& v) {
    m_world->m_batch->Queue(this, B2GPU_CMD_SET_LINEAR_VELOCITY, true, v[0], v[1], 0.0f, 0.0f);
}
inline void b2Body::SetAngularVelocity(float omega) {
    m_world->m_batch->Queue(this, B2GPU_CMD_SET_ANGULAR_VELOCITY, true, omega, 0.0f, 0.0f, 0.0f);
}
inline void b2Body::SetTransform(const b2Vec<float, 2>& position, float angle) {
    m_world->m_batch->Queue(this, B2GPU_CMD_SET_TRANSFORM, true, position[0], position[1], angle, 0.0f);
}
inline void b2Body::SetAwake(bool flag) {
    m_world->m_batch->Queue(this, B2GPU_CMD_SET_AWAKE, flag, 0.0f, 0.0f, 0.0f, 0.0f);
}
// The edits below only exist on the device: a scene that is still open is built first.
inline void b2WorldBatch::QueueEdit(const b2Body* b, int32_t op, int32_t arg, float vx, float vy, float px, float py) {
    BuildEarly();
    FlushSpawns();
    b2gpu_body_command c;
    c.world = b->m_world->m_index; c.body = b->m_index; c.op = op; c.wake = arg;
    c.v[0] = vx; c.v[1] = vy; c.p[0] = px; c.p[1] = py;
    m_commands.push_back(c);
}
inline void b2Body::SetType(b2BodyType type) {
    if (type == m_type) return;
    m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_TYPE, (int32_t)type, 0.0f, 0.0f, 0.0f, 0.0f);
    m_type = type;
}
inline void b2Body::SetActive(bool flag) { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_ACTIVE, flag ? 1 : 0, 0.0f, 0.0f, 0.0f, 0.0f); }
inline void b2Body::SetMassData(const b2MassData* d) {
    m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_MASS_DATA, 0, d->mass, d->I, d->center[0], d->center[1]);
}
inline void b2Body::ResetMassData() { m_world->m_batch->QueueEdit(this, B2GPU_CMD_RESET_MASS_DATA, 0, 0.0f, 0.0f, 0.0f, 0.0f); }
inline void b2Body::DestroyFixture(b2Fixture* fixture) {
    if (!fixture || fixture->GetBody() != this) return;
    m_world->m_batch->QueueEdit(this, B2GPU_CMD_DESTROY_FIXTURE, fixture->GetIndex(), 0.0f, 0.0f, 0.0f, 0.0f);
    fixture->m_body = nullptr;
}
inline void b2Body::QueueDamping() { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_DAMPING, 0, m_linearDamping, m_angularDamping, 0.0f, 0.0f); }
inline void b2Body::SetGravityScale(float scale) { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_GRAVITY_SCALE, 0, scale, 0.0f, 0.0f, 0.0f); }
inline void b2Body::SetBullet(bool flag) { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_BULLET, flag ? 1 : 0, 0.0f, 0.0f, 0.0f, 0.0f); }
inline void b2Body::SetSleepingAllowed(bool flag) { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_SLEEPING_ALLOWED, flag ? 1 : 0, 0.0f, 0.0f, 0.0f, 0.0f); }
inline void b2Body::SetFixedRotation(bool flag) { m_world->m_batch->QueueEdit(this, B2GPU_CMD_SET_FIXED_ROTATION, flag ? 1 : 0, 0.0f, 0.0f, 0.0f, 0.0f); }
inline void b2Fixture::SetSensor(bool sensor) {
    m_def.isSensor = sensor ? 1 : 0;
    if (m_body) m_body->m_world->m_batch->QueueEdit(m_body, B2GPU_CMD_SET_SENSOR, m_index, sensor ? 1.0f : 0.0f, 0.0f, 0.0f, 0.0f);
}
inline void b2Fixture::SetDensity(float density) {
    if (!m_body || !(density >= 0.0f)) return;  // b2Assert(b2IsValid(density) && density >= 0.0f)
    m_def.density = density;
    float md[4];
    b2WorldBatch::Check(b2gpu_host_fixture_mass_data(&m_def, m_defVertices.empty() ? nullptr : m_defVertices.data(),
                                                     (int32_t)(m_defVertices.size() / 2), md));
    m_body->m_world->m_batch->QueueEdit(m_body, B2GPU_CMD_SET_DENSITY, m_index, md[0], md[1], md[2], md[3]);
}
inline void b2Fixture::QueueMaterial() {
    if (m_body) m_body->m_world->m_batch->QueueEdit(m_body, B2GPU_CMD_SET_MATERIAL, m_index, m_friction, m_restitution, 0.0f, 0.0f);
}
inline void b2Fixture::SetFilterData(const b2Filter& filter) {
    if (!m_body) return;
    m_def.categoryBits = filter.categoryBits; m_def.maskBits = filter.maskBits; m_def.groupIndex = filter.groupIndex;
    const int32_t bits = (int32_t)((uint32_t)filter.categoryBits | ((uint32_t)filter.maskBits << 16)), group = filter.groupIndex;
    float v[2];
    std::memcpy(&v[0], &bits, 4);  // the two integers travel as bit patterns in the command's float slots
    std::memcpy(&v[1], &group, 4);
    b2WorldBatch* B = m_body->m_world->m_batch;
    B->QueueEdit(m_body, B2GPU_CMD_SET_FILTER, m_index, 0.0f, 0.0f, 0.0f, 0.0f);
    std::memcpy(B->m_commands.back().v, v, 8);
}

}  // namespace box2d

#endif
