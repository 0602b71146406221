// TEST INFRASTRUCTURE — not part of the product path.
//
// ref_harness.cpp: a C-ABI wrapper around the reference's own vendored, unmodified
// Box2D 2.3.2 ("vanilla", namespace box2dref, /root/reference/Regression/box2d-ref/Box2DRef).
// It is compiled by oracle/Makefile against the reference sources where they lie and linked
// with them into oracle/_ref/libb2ref.so.  No reference source is copied into this repo.
//
// What it gives the tests and bench.py:
//   * build a vanilla b2World from a b2g_scene (include/b2g_scene.h) through the public API
//     b2World::CreateBody / b2Body::CreateFixture (Dynamics/b2World.cpp:109, b2Body.cpp:168)
//   * b2World::Step (Dynamics/b2World.cpp:899) and dumps of body state, the world contact list in
//     list order (manifolds, point counts, ids), fat AABBs, and listener events
//   * direct calls of the narrow-phase functions (b2CollidePolygons, ... Collision/b2Collision.h)
//     and shape helpers (b2PolygonShape::Set / ComputeMass) for known-answer tests
//   * an all-core CPU baseline: one world per thread slot, T threads (bench.py --impl reference)
// The dumps need b2Fixture::m_proxies (the broad-phase proxy ids), which the library keeps private.
#define private public
#define protected public
#include <Box2DRef/Box2D.h>
#undef private
#undef protected

#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "../include/b2g_scene.h"

using namespace box2dref;

namespace {

struct Event {
    int32_t type;  // 0 begin, 1 end, 2 presolve, 3 postsolve
    int32_t fixtureA, childA, fixtureB, childB;
    int32_t count;
    float normalImpulses[2];
    float tangentImpulses[2];
};

struct Recorder : public b2ContactListener {
    std::vector<Event> events;
    bool enabled = false;
    void push(int type, b2Contact* c, const b2ContactImpulse* imp) {
        if (!enabled) return;
        Event e;
        memset(&e, 0, sizeof(e));
        e.type = type;
        e.fixtureA = (int32_t)(intptr_t)c->GetFixtureA()->GetUserData();
        e.fixtureB = (int32_t)(intptr_t)c->GetFixtureB()->GetUserData();
        e.childA = c->GetChildIndexA();
        e.childB = c->GetChildIndexB();
        if (imp) {
            e.count = imp->count;
            for (int i = 0; i < imp->count; ++i) {
                e.normalImpulses[i] = imp->normalImpulses[i];
                e.tangentImpulses[i] = imp->tangentImpulses[i];
            }
        } else {
            e.count = c->GetManifold()->pointCount;
        }
        events.push_back(e);
    }
    void BeginContact(b2Contact* c) override { push(0, c, nullptr); }
    void EndContact(b2Contact* c) override { push(1, c, nullptr); }
    void PreSolve(b2Contact* c, const b2Manifold*) override { push(2, c, nullptr); }
    void PostSolve(b2Contact* c, const b2ContactImpulse* imp) override { push(3, c, imp); }
};

struct RefWorld {
    b2World* world;
    std::vector<b2Body*> bodies;      // creation order
    std::vector<b2Fixture*> fixtures; // creation order (nullptr once the body was destroyed)
    std::vector<int> childCounts;     // proxies per fixture
    std::vector<b2Joint*> joints;     // creation order = joint ids (nullptr once destroyed, alone or with one of its bodies)
    int proxySlots = 0;               // high-water mark of proxy slots (see ref_world_get_proxies)
    Recorder recorder;
};

void build_polygon(b2PolygonShape& poly, const b2g_fixture_def& fd, const float* verts) {
    if (fd.shapeKind == B2G_POLY_BOX) {
        poly.SetAsBox(fd.p[0], fd.p[1]);
    } else if (fd.shapeKind == B2G_POLY_BOX_OFFSET) {
        poly.SetAsBox(fd.p[0], fd.p[1], b2Vec2(fd.center[0], fd.center[1]), fd.angle);
    } else {
        b2Vec2 pts[b2_maxPolygonVertices];
        int n = fd.vertexCount < b2_maxPolygonVertices ? fd.vertexCount : b2_maxPolygonVertices;
        for (int i = 0; i < n; ++i) pts[i].Set(verts[2 * (fd.firstVertex + i)], verts[2 * (fd.firstVertex + i) + 1]);
        poly.Set(pts, n);
    }
}

// b2World::CreateBody from a body definition; the body's index (creation order) is its userData
void add_body(RefWorld* rw, const b2g_body_def& d) {
    b2BodyDef bd;
    bd.type = (b2BodyType)d.type;
    bd.position.Set(d.position[0], d.position[1]);
    bd.angle = d.angle;
    bd.linearVelocity.Set(d.linearVelocity[0], d.linearVelocity[1]);
    bd.angularVelocity = d.angularVelocity;
    bd.linearDamping = d.linearDamping;
    bd.angularDamping = d.angularDamping;
    bd.gravityScale = d.gravityScale;
    bd.allowSleep = (d.flags & B2G_BD_ALLOW_SLEEP) != 0;
    bd.awake = (d.flags & B2G_BD_AWAKE) != 0;
    bd.fixedRotation = (d.flags & B2G_BD_FIXED_ROTATION) != 0;
    bd.bullet = (d.flags & B2G_BD_BULLET) != 0;
    bd.active = (d.flags & B2G_BD_ACTIVE) != 0;
    bd.userData = (void*)(intptr_t)rw->bodies.size();
    rw->bodies.push_back(rw->world->CreateBody(&bd));
}

void note_proxy_slots(RefWorld* rw, b2Fixture* f) {
    for (int k = 0; k < f->m_proxyCount; ++k) {
        const int id = f->m_proxies[k].proxyId;
        const int slot = id == 0 ? 0 : (id + 1) / 2;
        if (slot + 1 > rw->proxySlots) rw->proxySlots = slot + 1;
    }
}

// b2Body::CreateFixture from a fixture definition; the fixture's index (creation order) is its userData
void add_fixture(RefWorld* rw, b2Body* body, const b2g_fixture_def& d, const float* v) {
    b2FixtureDef fd;
    fd.friction = d.friction;
    fd.restitution = d.restitution;
    fd.density = d.density;
    fd.isSensor = d.isSensor != 0;
    fd.filter.categoryBits = d.categoryBits;
    fd.filter.maskBits = d.maskBits;
    fd.filter.groupIndex = d.groupIndex;
    fd.userData = (void*)(intptr_t)rw->fixtures.size();
    b2Fixture* f = nullptr;
    switch (d.shapeType) {
        case B2G_SHAPE_CIRCLE: {
            b2CircleShape sh;
            sh.m_p.Set(d.p[0], d.p[1]);
            sh.m_radius = d.radius;
            fd.shape = &sh;
            f = body->CreateFixture(&fd);
        } break;
        case B2G_SHAPE_EDGE: {
            b2EdgeShape sh;
            sh.Set(b2Vec2(v[2 * d.firstVertex], v[2 * d.firstVertex + 1]), b2Vec2(v[2 * d.firstVertex + 2], v[2 * d.firstVertex + 3]));
            if (d.hasVertex0) { sh.m_vertex0.Set(d.vertex0[0], d.vertex0[1]); sh.m_hasVertex0 = true; }
            if (d.hasVertex3) { sh.m_vertex3.Set(d.vertex3[0], d.vertex3[1]); sh.m_hasVertex3 = true; }
            fd.shape = &sh;
            f = body->CreateFixture(&fd);
        } break;
        case B2G_SHAPE_POLYGON: {
            b2PolygonShape sh;
            build_polygon(sh, d, v);
            fd.shape = &sh;
            f = body->CreateFixture(&fd);
        } break;
        case B2G_SHAPE_CHAIN: {
            b2ChainShape sh;
            std::vector<b2Vec2> pts(d.vertexCount);
            for (int k = 0; k < d.vertexCount; ++k) pts[k].Set(v[2 * (d.firstVertex + k)], v[2 * (d.firstVertex + k) + 1]);
            if (d.shapeKind == B2G_CHAIN_LOOP) {
                sh.CreateLoop(pts.data(), d.vertexCount);
            } else {
                sh.CreateChain(pts.data(), d.vertexCount);
                if (d.hasVertex0) sh.SetPrevVertex(b2Vec2(d.vertex0[0], d.vertex0[1]));
                if (d.hasVertex3) sh.SetNextVertex(b2Vec2(d.vertex3[0], d.vertex3[1]));
            }
            fd.shape = &sh;
            f = body->CreateFixture(&fd);
        } break;
    }
    rw->fixtures.push_back(f);
    rw->childCounts.push_back(f->GetShape()->GetChildCount());
    note_proxy_slots(rw, f);
}

// b2World::CreateJoint from a joint definition; the joint's id is its place in rw->joints (creation order)
b2Joint* add_joint(RefWorld* rw, const b2g_joint_def& d) {
    b2World* w = rw->world;
    b2Joint* made = nullptr;
    {
        b2Body* bA = rw->bodies[d.bodyA];
        b2Body* bB = rw->bodies[d.bodyB];
        if (d.type == B2G_JOINT_MOUSE) {
            b2MouseJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.target.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.maxForce = d.length; jd.frequencyHz = d.frequencyHz; jd.dampingRatio = d.dampingRatio;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_GEAR) {
            auto jointAt = [&](int k) { return rw->joints[k]; };  // joints[k] of the scene = the k-th CreateJoint call
            b2GearJointDef jd;
            jd.joint1 = jointAt(d.enableLimit);
            jd.joint2 = jointAt(d.enableMotor);
            jd.bodyA = jd.joint1->GetBodyB(); jd.bodyB = jd.joint2->GetBodyB();  // what the Testbed passes; the constructor overrides them
            jd.collideConnected = d.collideConnected != 0;
            jd.ratio = d.motorSpeed;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_REVOLUTE) {
            b2RevoluteJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.referenceAngle = d.referenceAngle;
            jd.enableLimit = d.enableLimit != 0; jd.lowerAngle = d.lower; jd.upperAngle = d.upper;
            jd.enableMotor = d.enableMotor != 0; jd.maxMotorTorque = d.maxMotor; jd.motorSpeed = d.motorSpeed;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_PRISMATIC) {
            b2PrismaticJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.localAxisA.Set(d.localAxisA[0], d.localAxisA[1]);
            jd.referenceAngle = d.referenceAngle;
            jd.enableLimit = d.enableLimit != 0; jd.lowerTranslation = d.lower; jd.upperTranslation = d.upper;
            jd.enableMotor = d.enableMotor != 0; jd.maxMotorForce = d.maxMotor; jd.motorSpeed = d.motorSpeed;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_WHEEL) {
            b2WheelJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.localAxisA.Set(d.localAxisA[0], d.localAxisA[1]);
            jd.enableMotor = d.enableMotor != 0; jd.maxMotorTorque = d.maxMotor; jd.motorSpeed = d.motorSpeed;
            jd.frequencyHz = d.frequencyHz; jd.dampingRatio = d.dampingRatio;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_ROPE) {
            b2RopeJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.maxLength = d.length;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_PULLEY) {
            b2PulleyJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.groundAnchorA.Set(d.localAxisA[0], d.localAxisA[1]);
            jd.groundAnchorB.Set(d.lower, d.upper);
            jd.lengthA = d.length; jd.lengthB = d.maxMotor; jd.ratio = d.motorSpeed;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_MOTOR) {
            b2MotorJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.linearOffset.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.angularOffset = d.referenceAngle;
            jd.maxForce = d.length; jd.maxTorque = d.maxMotor; jd.correctionFactor = d.frequencyHz;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_FRICTION) {
            b2FrictionJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.maxForce = d.length; jd.maxTorque = d.maxMotor;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_WELD) {
            b2WeldJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.referenceAngle = d.referenceAngle;
            jd.frequencyHz = d.frequencyHz; jd.dampingRatio = d.dampingRatio;
            made = w->CreateJoint(&jd);
        } else if (d.type == B2G_JOINT_DISTANCE) {
            b2DistanceJointDef jd;
            jd.bodyA = bA; jd.bodyB = bB; jd.collideConnected = d.collideConnected != 0;
            jd.localAnchorA.Set(d.localAnchorA[0], d.localAnchorA[1]);
            jd.localAnchorB.Set(d.localAnchorB[0], d.localAnchorB[1]);
            jd.length = d.length; jd.frequencyHz = d.frequencyHz; jd.dampingRatio = d.dampingRatio;
            made = w->CreateJoint(&jd);
        }
    }
    rw->joints.push_back(made);
    return made;
}

RefWorld* build_world(const b2g_scene* s) {
    RefWorld* rw = new RefWorld;
    rw->world = new b2World(b2Vec2(s->gravity[0], s->gravity[1]));
    b2World* w = rw->world;
    w->SetAllowSleeping((s->worldFlags & B2G_WORLD_ALLOW_SLEEP) != 0);
    w->SetWarmStarting((s->worldFlags & B2G_WORLD_WARM_STARTING) != 0);
    w->SetContinuousPhysics((s->worldFlags & B2G_WORLD_CONTINUOUS) != 0);
    w->SetAutoClearForces((s->worldFlags & B2G_WORLD_CLEAR_FORCES) != 0);
    w->SetContactListener(&rw->recorder);

    for (int i = 0; i < s->bodyCount; ++i) add_body(rw, s->bodies[i]);
    for (int i = 0; i < s->fixtureCount; ++i) add_fixture(rw, rw->bodies[s->fixtures[i].body], s->fixtures[i], s->vertices);
    // b2World::CreateJoint in call order, from the same local-frame definitions the device path receives
    for (int i = 0; i < s->jointCount; ++i) add_joint(rw, s->joints[i]);
    return rw;
}

}  // namespace

extern "C" {

void* ref_world_create(const b2g_scene* s) { return build_world(s); }

void ref_world_destroy(void* h) {
    RefWorld* rw = (RefWorld*)h;
    delete rw->world;
    delete rw;
}

void ref_world_record_events(void* h, int on) { ((RefWorld*)h)->recorder.enabled = on != 0; }

void ref_world_step(void* h, float dt, int velIters, int posIters) {
    RefWorld* rw = (RefWorld*)h;
    rw->recorder.events.clear();
    rw->world->Step(dt, velIters, posIters);
}

// b2World::CreateBody + b2Body::CreateFixture after the world has been stepped; returns the new body's index.
int ref_world_create_body(void* h, const b2g_body_def* bd, const b2g_fixture_def* fds, int nf, const float* vertices) {
    RefWorld* rw = (RefWorld*)h;
    add_body(rw, *bd);
    b2Body* body = rw->bodies.back();
    for (int i = 0; i < nf; ++i) add_fixture(rw, body, fds[i], vertices);
    return (int)rw->bodies.size() - 1;
}

// b2Fixture::SetFilterData (the B2GPU_CMD_SET_FILTER payload as plain integers)
void ref_world_set_filter(void* h, int fixture, int categoryBits, int maskBits, int groupIndex) {
    RefWorld* rw = (RefWorld*)h;
    if (fixture < 0 || fixture >= (int)rw->fixtures.size() || !rw->fixtures[fixture]) return;
    b2Filter f;
    f.categoryBits = (uint16)categoryBits; f.maskBits = (uint16)maskBits; f.groupIndex = (int16)groupIndex;
    rw->fixtures[fixture]->SetFilterData(f);
}

// b2Fixture::SetDensity (b2Fixture.h:313-317): the body's mass data waits for b2Body::ResetMassData (op 16)
void ref_world_set_density(void* h, int fixture, float density) {
    RefWorld* rw = (RefWorld*)h;
    if (fixture < 0 || fixture >= (int)rw->fixtures.size() || !rw->fixtures[fixture]) return;
    rw->fixtures[fixture]->SetDensity(density);
}

// b2Fixture::GetMassData: out = mass, I, center.x, center.y (the B2GPU_CMD_SET_DENSITY payload)
int ref_world_fixture_mass_data(void* h, int fixture, float* out) {
    RefWorld* rw = (RefWorld*)h;
    if (fixture < 0 || fixture >= (int)rw->fixtures.size() || !rw->fixtures[fixture]) return -1;
    b2MassData md;
    rw->fixtures[fixture]->GetMassData(&md);
    out[0] = md.mass; out[1] = md.I; out[2] = md.center.x; out[3] = md.center.y;
    return 0;
}

// b2World::CreateJoint on a stepped world; returns the joint id.
int ref_world_create_joint(void* h, const b2g_joint_def* d) {
    RefWorld* rw = (RefWorld*)h;
    if (d->bodyA < 0 || d->bodyA >= (int)rw->bodies.size() || d->bodyB < 0 || d->bodyB >= (int)rw->bodies.size()) return -1;
    if (!rw->bodies[d->bodyA] || !rw->bodies[d->bodyB]) return -1;
    if (d->type == B2G_JOINT_GEAR)  // joint1 / joint2 by joint id: they must still exist
        for (int k : {d->enableLimit, d->enableMotor})
            if (k < 0 || k >= (int)rw->joints.size() || !rw->joints[k]) return -1;
    add_joint(rw, *d);
    return (int)rw->joints.size() - 1;
}

// b2Body::CreateFixture on an existing body of a stepped world; returns the fixture index.
int ref_world_create_fixture(void* h, int body, const b2g_fixture_def* fd, const float* vertices) {
    RefWorld* rw = (RefWorld*)h;
    if (body < 0 || body >= (int)rw->bodies.size() || !rw->bodies[body]) return -1;
    add_fixture(rw, rw->bodies[body], *fd, vertices);
    return (int)rw->fixtures.size() - 1;
}

// Between-step body edits, ops = B2GPU_CMD_* of include/b2gpu.h: the call a user of the reference makes on a b2Body.
void ref_world_body_command(void* h, int body, int op, int wake, float vx, float vy, float px, float py) {
    RefWorld* rw = (RefWorld*)h;
    if (op == 12) {  // b2World::DestroyJoint(joint id)
        if (wake < 0 || wake >= (int)rw->joints.size() || !rw->joints[wake]) return;
        rw->world->DestroyJoint(rw->joints[wake]);
        rw->joints[wake] = nullptr;
        return;
    }
    if (body < 0 || body >= (int)rw->bodies.size()) return;
    b2Body* b = rw->bodies[body];
    if (!b) return;  // destroyed earlier
    if (op >= 11) {
        switch (op) {
            case 11: {  // b2Fixture::SetFilterData: the two ints travel as float bit patterns
                if (wake < 0 || wake >= (int)rw->fixtures.size() || !rw->fixtures[wake]) return;
                int32_t bits, group;
                memcpy(&bits, &vx, 4); memcpy(&group, &vy, 4);
                b2Filter f;
                f.categoryBits = (uint16)(bits & 0xffff); f.maskBits = (uint16)((bits >> 16) & 0xffff); f.groupIndex = (int16)group;
                rw->fixtures[wake]->SetFilterData(f);
            } break;
            case 13: b->SetType((b2BodyType)wake); break;
            case 14: {
                b->SetActive(wake != 0);
                for (b2Fixture* f : rw->fixtures) if (f && f->GetBody() == b) note_proxy_slots(rw, f);
            } break;
            case 15: { b2MassData md; md.mass = vx; md.I = vy; md.center.Set(px, py); b->SetMassData(&md); } break;
            case 16: b->ResetMassData(); break;
            case 17: {
                if (wake < 0 || wake >= (int)rw->fixtures.size() || !rw->fixtures[wake] || rw->fixtures[wake]->GetBody() != b) return;
                b->DestroyFixture(rw->fixtures[wake]);
                rw->fixtures[wake] = nullptr;
            } break;
            case 18: b->SetLinearDamping(vx); b->SetAngularDamping(vy); break;
            case 19: b->SetGravityScale(vx); break;
            case 20: b->SetBullet(wake != 0); break;
            case 21: b->SetSleepingAllowed(wake != 0); break;
            case 22: b->SetFixedRotation(wake != 0); break;
            case 23:
            case 24: {
                if (wake < 0 || wake >= (int)rw->fixtures.size() || !rw->fixtures[wake] || rw->fixtures[wake]->GetBody() != b) return;
                if (op == 23) rw->fixtures[wake]->SetSensor(vx != 0.0f);
                else { rw->fixtures[wake]->SetFriction(vx); rw->fixtures[wake]->SetRestitution(vy); }
            } break;
        }
        return;
    }
    if (op == 9) {   // b2World::DestroyBody: the body's fixtures and joints go with it; the indices stay reserved
        for (size_t i = 0; i < rw->fixtures.size(); ++i)
            if (rw->fixtures[i] && rw->fixtures[i]->GetBody() == b) rw->fixtures[i] = nullptr;
        for (size_t i = 0; i < rw->joints.size(); ++i)
            if (rw->joints[i] && (rw->joints[i]->GetBodyA() == b || rw->joints[i]->GetBodyB() == b)) rw->joints[i] = nullptr;
        rw->world->DestroyBody(b);
        rw->bodies[body] = nullptr;
        return;
    }
    if (op == 10) {  // b2MouseJoint::SetTarget on the newest mouse joint whose bodyB is this body
        for (b2JointEdge* je = b->GetJointList(); je; je = je->next) {
            if (je->joint->GetType() == e_mouseJoint && je->joint->GetBodyB() == b) {
                static_cast<b2MouseJoint*>(je->joint)->SetTarget(b2Vec2(vx, vy));
                break;
            }
        }
        return;
    }
    switch (op) {
        case 0: b->ApplyForce(b2Vec2(vx, vy), b2Vec2(px, py), wake != 0); break;
        case 1: b->ApplyForceToCenter(b2Vec2(vx, vy), wake != 0); break;
        case 2: b->ApplyTorque(vx, wake != 0); break;
        case 3: b->ApplyLinearImpulse(b2Vec2(vx, vy), b2Vec2(px, py), wake != 0); break;
        case 4: b->ApplyAngularImpulse(vx, wake != 0); break;
        case 5: b->SetLinearVelocity(b2Vec2(vx, vy)); break;
        case 6: b->SetAngularVelocity(vx); break;
        case 7: b->SetTransform(b2Vec2(vx, vy), px); break;
        case 8: b->SetAwake(wake != 0); break;
    }
}

// b2World::QueryAABB: (fixture index, child) of every reported proxy, in the reference's (tree) order.
namespace {
struct QueryCollector : public b2QueryCallback {
    std::vector<int32_t> out;
    bool ReportFixture(b2Fixture*) override { return true; }
};
}  // namespace
int ref_world_query_aabb(void* h, const float* lower, const float* upper, int32_t* out, int cap) {
    // b2QueryCallback only hands over the fixture; the child index lives in the proxy, so go through the broad
    // phase like b2World::QueryAABB's wrapper does (Dynamics/b2World.cpp:965-985).
    RefWorld* rw = (RefWorld*)h;
    struct Wrapper {
        const b2BroadPhase* bp;
        std::vector<int32_t>* out;
        bool QueryCallback(int32 proxyId) {
            b2FixtureProxy* proxy = (b2FixtureProxy*)bp->GetUserData(proxyId);
            out->push_back((int32_t)(intptr_t)proxy->fixture->GetUserData() | (proxy->childIndex << 16));
            return true;
        }
    };
    std::vector<int32_t> hits;
    Wrapper wrap{&rw->world->GetContactManager().m_broadPhase, &hits};
    b2AABB aabb;
    aabb.lowerBound.Set(lower[0], lower[1]);
    aabb.upperBound.Set(upper[0], upper[1]);
    wrap.bp->Query(&wrap, aabb);
    for (int i = 0; i < (int)hits.size() && i < cap; ++i) out[i] = hits[i];
    return (int)hits.size();
}

// b2World::RayCast with the closest-hit callback.  out: fixture, child | fraction, point.xy, normal.xy
namespace {
struct ClosestHit : public b2RayCastCallback {
    b2Fixture* fixture = nullptr;
    b2Vec2 point, normal;
    float32 fraction = 0.0f;
    float32 ReportFixture(b2Fixture* f, const b2Vec2& p, const b2Vec2& n, float32 fr) override {
        fixture = f; point = p; normal = n; fraction = fr;
        return fr;
    }
};
}  // namespace
int ref_world_raycast_closest(void* h, const float* p1, const float* p2, int32_t* outI, float* outF) {
    RefWorld* rw = (RefWorld*)h;
    ClosestHit cb;
    rw->world->RayCast(&cb, b2Vec2(p1[0], p1[1]), b2Vec2(p2[0], p2[1]));
    if (!cb.fixture) { outI[0] = -1; outI[1] = 0; return 0; }
    outI[0] = (int32_t)(intptr_t)cb.fixture->GetUserData();
    // the child index is not reported to the callback: recover it by re-casting the fixture's children
    outI[1] = 0;
    b2RayCastInput in; in.p1.Set(p1[0], p1[1]); in.p2.Set(p2[0], p2[1]); in.maxFraction = 1.0f;
    for (int32 c = 0; c < cb.fixture->GetShape()->GetChildCount(); ++c) {
        b2RayCastOutput o;
        if (cb.fixture->RayCast(&o, in, c) && o.fraction == cb.fraction && o.normal.x == cb.normal.x && o.normal.y == cb.normal.y) {
            outI[1] = c;
            break;
        }
    }
    outF[0] = cb.fraction; outF[1] = cb.point.x; outF[2] = cb.point.y; outF[3] = cb.normal.x; outF[4] = cb.normal.y;
    return 1;
}

int ref_world_body_count(void* h) { return (int)((RefWorld*)h)->bodies.size(); }
int ref_world_contact_count(void* h) { return ((RefWorld*)h)->world->GetContactCount(); }
int ref_world_proxy_count(void* h) { return ((RefWorld*)h)->world->GetProxyCount(); }

// 16 floats per body, creation order:
// xf.p.x xf.p.y xf.q.s xf.q.c | sweep.c.x c.y a | v.x v.y w | awake sleepTime(n/a=0) | mass invMass I(about COM+..GetInertia) localCenter.x localCenter.y  (15) + pad
void ref_world_get_bodies(void* h, float* out) {
    RefWorld* rw = (RefWorld*)h;
    for (size_t i = 0; i < rw->bodies.size(); ++i) {
        b2Body* b = rw->bodies[i];
        float* o = out + 16 * i;
        if (!b) {  // destroyed: a row of zeros with the marker in the last column
            for (int k = 0; k < 15; ++k) o[k] = 0.0f;
            o[15] = 1.0f;
            continue;
        }
        const b2Transform& xf = b->GetTransform();
        o[0] = xf.p.x; o[1] = xf.p.y; o[2] = xf.q.s; o[3] = xf.q.c;
        o[4] = b->GetWorldCenter().x; o[5] = b->GetWorldCenter().y; o[6] = b->GetAngle();
        o[7] = b->GetLinearVelocity().x; o[8] = b->GetLinearVelocity().y; o[9] = b->GetAngularVelocity();
        o[10] = b->IsAwake() ? 1.0f : 0.0f;
        o[11] = b->GetMass();
        o[12] = b->GetInertia();
        o[13] = b->GetLocalCenter().x; o[14] = b->GetLocalCenter().y;
        o[15] = 0.0f;
    }
}

// Contacts in world-list order (head first).  ints: 8 per contact
//   fixtureA childA fixtureB childB flags(bit0 touching, bit1 enabled) pointCount key0 key1
// floats: 16 per contact: localNormal(2) localPoint(2) type p0.local(2) p0.nI p0.tI p1.local(2) p1.nI p1.tI friction restitution pad
int ref_world_get_contacts(void* h, int32_t* ints, float* floats, int cap) {
    RefWorld* rw = (RefWorld*)h;
    int n = 0;
    for (b2Contact* c = rw->world->GetContactList(); c; c = c->GetNext()) {
        if (n < cap) {
            int32_t* I = ints + 8 * n;
            float* F = floats + 16 * n;
            const b2Manifold* m = c->GetManifold();
            I[0] = (int32_t)(intptr_t)c->GetFixtureA()->GetUserData();
            I[1] = c->GetChildIndexA();
            I[2] = (int32_t)(intptr_t)c->GetFixtureB()->GetUserData();
            I[3] = c->GetChildIndexB();
            I[4] = (c->IsTouching() ? 1 : 0) | (c->IsEnabled() ? 2 : 0);
            I[5] = m->pointCount;
            I[6] = m->pointCount > 0 ? (int32_t)m->points[0].id.key : 0;
            I[7] = m->pointCount > 1 ? (int32_t)m->points[1].id.key : 0;
            memset(F, 0, 16 * sizeof(float));
            if (m->pointCount > 0) {
                F[0] = m->localNormal.x; F[1] = m->localNormal.y;
                F[2] = m->localPoint.x; F[3] = m->localPoint.y;
                F[4] = (float)m->type;
                for (int k = 0; k < m->pointCount; ++k) {
                    F[5 + 4 * k] = m->points[k].localPoint.x;
                    F[6 + 4 * k] = m->points[k].localPoint.y;
                    F[7 + 4 * k] = m->points[k].normalImpulse;
                    F[8 + 4 * k] = m->points[k].tangentImpulse;
                }
            }
            F[13] = c->GetFriction();
            F[14] = c->GetRestitution();
        }
        ++n;
    }
    return n;
}

// Fat AABBs per proxy: for each fixture (creation order), each child: 4 floats + proxyId.
int ref_world_get_proxies(void* h, float* aabbs, int32_t* ids, int cap) {
    // One row per proxy SLOT: slot k holds the proxy whose tree id is 0 (k = 0) or 2k - 1 (a leaf
    // and its parent are allocated together, so leaf ids are the odd numbers; a destroyed proxy's id is reused by the
    // next CreateProxy).  Rows of slots without a live proxy are zero with id -1.
    RefWorld* rw = (RefWorld*)h;
    const b2BroadPhase& bp = rw->world->GetContactManager().m_broadPhase;
    int n = rw->proxySlots;
    for (int i = 0; i < n && i < cap; ++i) { aabbs[4 * i] = aabbs[4 * i + 1] = aabbs[4 * i + 2] = aabbs[4 * i + 3] = 0.0f; ids[i] = -1; }
    for (b2Fixture* f : rw->fixtures) {
        if (!f) continue;
        for (int k = 0; k < f->m_proxyCount; ++k) {
            const int id = f->m_proxies[k].proxyId;
            const int slot = id == 0 ? 0 : (id + 1) / 2;
            if (slot >= cap) continue;
            const b2AABB& a = bp.GetFatAABB(id);
            aabbs[4 * slot + 0] = a.lowerBound.x; aabbs[4 * slot + 1] = a.lowerBound.y;
            aabbs[4 * slot + 2] = a.upperBound.x; aabbs[4 * slot + 3] = a.upperBound.y;
            ids[slot] = id;
        }
    }
    return n;
}

int ref_world_get_events(void* h, int32_t* ints, float* floats, int cap) {
    RefWorld* rw = (RefWorld*)h;
    int n = (int)rw->recorder.events.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const Event& e = rw->recorder.events[i];
        int32_t* I = ints + 6 * i;
        I[0] = e.type; I[1] = e.fixtureA; I[2] = e.childA; I[3] = e.fixtureB; I[4] = e.childB; I[5] = e.count;
        float* F = floats + 4 * i;
        F[0] = e.normalImpulses[0]; F[1] = e.normalImpulses[1];
        F[2] = e.tangentImpulses[0]; F[3] = e.tangentImpulses[1];
    }
    return n;
}

// ---- stateless known-answer entry points -------------------------------------------------------

// Polygon after b2PolygonShape::Set / SetAsBox: out = count, then 8 verts, 8 normals, centroid (2+32+2 floats)
int ref_polygon_build(const b2g_fixture_def* fd, const float* verts, float* out) {
    b2PolygonShape p;
    build_polygon(p, *fd, verts);
    for (int i = 0; i < 8; ++i) {
        out[2 * i] = i < p.m_count ? p.m_vertices[i].x : 0.0f;
        out[2 * i + 1] = i < p.m_count ? p.m_vertices[i].y : 0.0f;
        out[16 + 2 * i] = i < p.m_count ? p.m_normals[i].x : 0.0f;
        out[16 + 2 * i + 1] = i < p.m_count ? p.m_normals[i].y : 0.0f;
    }
    out[32] = p.m_centroid.x; out[33] = p.m_centroid.y;
    b2MassData md;
    p.ComputeMass(&md, fd->density);
    out[34] = md.mass; out[35] = md.center.x; out[36] = md.center.y; out[37] = md.I;
    return p.m_count;
}

static void fill_poly(b2PolygonShape& p, const float* poly /* count, 8v, 8n, centroid */) {
    p.m_count = (int)poly[0];
    for (int i = 0; i < 8; ++i) {
        p.m_vertices[i].Set(poly[1 + 2 * i], poly[2 + 2 * i]);
        p.m_normals[i].Set(poly[17 + 2 * i], poly[18 + 2 * i]);
    }
    p.m_centroid.Set(poly[33], poly[34]);
}

static void dump_manifold(const b2Manifold& m, int32_t* ints, float* floats) {
    ints[0] = m.pointCount;
    ints[1] = (int)m.type;
    ints[2] = m.pointCount > 0 ? (int32_t)m.points[0].id.key : 0;
    ints[3] = m.pointCount > 1 ? (int32_t)m.points[1].id.key : 0;
    memset(floats, 0, 8 * sizeof(float));
    if (m.pointCount > 0) {
        floats[0] = m.localNormal.x; floats[1] = m.localNormal.y;
        floats[2] = m.localPoint.x; floats[3] = m.localPoint.y;
        for (int k = 0; k < m.pointCount; ++k) {
            floats[4 + 2 * k] = m.points[k].localPoint.x;
            floats[5 + 2 * k] = m.points[k].localPoint.y;
        }
    }
}

static b2Transform mk_xf(const float* x) {
    b2Transform xf;
    xf.p.Set(x[0], x[1]);
    xf.q.s = x[2];
    xf.q.c = x[3];
    return xf;
}

// polys: 35 floats each (count, 8 verts, 8 normals, centroid); xf: px py s c
void ref_collide_polygons(const float* polyA, const float* xfA, const float* polyB, const float* xfB,
                          int32_t* ints, float* floats) {
    b2PolygonShape a, b;
    fill_poly(a, polyA);
    fill_poly(b, polyB);
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    b2CollidePolygons(&m, &a, mk_xf(xfA), &b, mk_xf(xfB));
    dump_manifold(m, ints, floats);
}

void ref_collide_circles(const float* cA /* px py r */, const float* xfA, const float* cB, const float* xfB,
                         int32_t* ints, float* floats) {
    b2CircleShape a, b;
    a.m_p.Set(cA[0], cA[1]); a.m_radius = cA[2];
    b.m_p.Set(cB[0], cB[1]); b.m_radius = cB[2];
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    b2CollideCircles(&m, &a, mk_xf(xfA), &b, mk_xf(xfB));
    dump_manifold(m, ints, floats);
}

void ref_collide_polygon_circle(const float* polyA, const float* xfA, const float* cB, const float* xfB,
                                int32_t* ints, float* floats) {
    b2PolygonShape a;
    fill_poly(a, polyA);
    b2CircleShape b;
    b.m_p.Set(cB[0], cB[1]); b.m_radius = cB[2];
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    b2CollidePolygonAndCircle(&m, &a, mk_xf(xfA), &b, mk_xf(xfB));
    dump_manifold(m, ints, floats);
}

static void fill_edge(b2EdgeShape& e, const float* ed /* v0 v1 v2 v3 has0 has3 */) {
    e.m_vertex0.Set(ed[0], ed[1]);
    e.m_vertex1.Set(ed[2], ed[3]);
    e.m_vertex2.Set(ed[4], ed[5]);
    e.m_vertex3.Set(ed[6], ed[7]);
    e.m_hasVertex0 = ed[8] != 0.0f;
    e.m_hasVertex3 = ed[9] != 0.0f;
}

void ref_collide_edge_circle(const float* edge, const float* xfA, const float* cB, const float* xfB,
                             int32_t* ints, float* floats) {
    b2EdgeShape a;
    fill_edge(a, edge);
    b2CircleShape b;
    b.m_p.Set(cB[0], cB[1]); b.m_radius = cB[2];
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    b2CollideEdgeAndCircle(&m, &a, mk_xf(xfA), &b, mk_xf(xfB));
    dump_manifold(m, ints, floats);
}

void ref_collide_edge_polygon(const float* edge, const float* xfA, const float* polyB, const float* xfB,
                              int32_t* ints, float* floats) {
    b2EdgeShape a;
    fill_edge(a, edge);
    b2PolygonShape b;
    fill_poly(b, polyB);
    b2Manifold m;
    memset(&m, 0, sizeof(m));
    b2CollideEdgeAndPolygon(&m, &a, mk_xf(xfA), &b, mk_xf(xfB));
    dump_manifold(m, ints, floats);
}

// libm as the reference links it (b2Rot::Set, Common/b2Math.h:302-311)
void ref_sincosf(const float* x, float* s, float* c, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { s[i] = sinf(x[i]); c[i] = cosf(x[i]); }
}

// ---- all-core CPU baseline ---------------------------------------------------------------------
// One world per slot, worldsPerThread slots per thread, T threads pinned to cores; only the step
// loop is timed (steady_clock), after all threads have built their worlds.
// Returns wall seconds between the common start and the last thread's finish.
double ref_bench_run(const b2g_scene* scenes, int sceneCount, int worldsPerThread, int threads, int steps,
                     int warmupSteps, float dt, int velIters, int posIters, uint64_t* checksum) {
    // Touch the library's lazily-initialised globals once on this thread.
    {
        RefWorld* rw = build_world(&scenes[0]);
        rw->world->Step(dt, velIters, posIters);
        delete rw->world;
        delete rw;
    }
    std::atomic<int> ready(0), go(0);
    std::vector<double> elapsed(threads, 0.0);
    std::vector<uint64_t> sums(threads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            cpu_set_t set;
            CPU_ZERO(&set);
            CPU_SET(t % (int)std::thread::hardware_concurrency(), &set);
            pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
            std::vector<RefWorld*> worlds;
            for (int k = 0; k < worldsPerThread; ++k)
                worlds.push_back(build_world(&scenes[(t * worldsPerThread + k) % sceneCount]));
            for (int s = 0; s < warmupSteps; ++s)
                for (RefWorld* rw : worlds) rw->world->Step(dt, velIters, posIters);
            ready.fetch_add(1);
            while (go.load() == 0) std::this_thread::yield();
            auto t0 = std::chrono::steady_clock::now();
            for (int s = 0; s < steps; ++s)
                for (RefWorld* rw : worlds) rw->world->Step(dt, velIters, posIters);
            auto t1 = std::chrono::steady_clock::now();
            elapsed[t] = std::chrono::duration<double>(t1 - t0).count();
            uint64_t acc = 0;
            for (RefWorld* rw : worlds) {
                for (b2Body* b : rw->bodies) {
                    float y = b->GetPosition().y;
                    uint32_t u;
                    memcpy(&u, &y, 4);
                    acc = acc * 1099511628211ull + u;
                }
                delete rw->world;
                delete rw;
            }
            sums[t] = acc;
        });
    }
    while (ready.load() < threads) std::this_thread::yield();
    go.store(1);
    for (auto& th : pool) th.join();
    double worst = 0.0;
    uint64_t acc = 0;
    for (int t = 0; t < threads; ++t) {
        if (elapsed[t] > worst) worst = elapsed[t];
        acc ^= sums[t];
    }
    if (checksum) *checksum = acc;
    return worst;
}

// The same with the timed region repeated: `repeats` back-to-back blocks of `steps` steps, all threads released together
// at the start of each block; seconds[r] = the slowest thread's time for block r.  (bench.py --impl reference reports the
// median and the spread, so that a ratio against it carries its own error bar.)
double ref_bench_run_repeats(const b2g_scene* scenes, int sceneCount, int worldsPerThread, int threads, int steps,
                             int warmupSteps, float dt, int velIters, int posIters, int repeats, double* seconds) {
    {
        RefWorld* rw = build_world(&scenes[0]);
        rw->world->Step(dt, velIters, posIters);
        delete rw->world;
        delete rw;
    }
    if (repeats < 1) repeats = 1;
    std::atomic<int> arrived(0), released(0);
    std::vector<double> elapsed((size_t)threads * repeats, 0.0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            cpu_set_t set;
            CPU_ZERO(&set);
            CPU_SET(t % (int)std::thread::hardware_concurrency(), &set);
            pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
            std::vector<RefWorld*> worlds;
            for (int k = 0; k < worldsPerThread; ++k)
                worlds.push_back(build_world(&scenes[(t * worldsPerThread + k) % sceneCount]));
            for (int s = 0; s < warmupSteps; ++s)
                for (RefWorld* rw : worlds) rw->world->Step(dt, velIters, posIters);
            for (int r = 0; r < repeats; ++r) {
                arrived.fetch_add(1);
                while (released.load() <= r) std::this_thread::yield();
                auto t0 = std::chrono::steady_clock::now();
                for (int s = 0; s < steps; ++s)
                    for (RefWorld* rw : worlds) rw->world->Step(dt, velIters, posIters);
                auto t1 = std::chrono::steady_clock::now();
                elapsed[(size_t)t * repeats + r] = std::chrono::duration<double>(t1 - t0).count();
            }
            for (RefWorld* rw : worlds) {
                delete rw->world;
                delete rw;
            }
        });
    }
    for (int r = 0; r < repeats; ++r) {
        while (arrived.load() < threads * (r + 1)) std::this_thread::yield();
        released.store(r + 1);
    }
    for (auto& th : pool) th.join();
    double best = 0.0;
    for (int r = 0; r < repeats; ++r) {
        double worst = 0.0;
        for (int t = 0; t < threads; ++t) worst = elapsed[(size_t)t * repeats + r] > worst ? elapsed[(size_t)t * repeats + r] : worst;
        if (seconds) seconds[r] = worst;
        if (r == 0 || worst < best) best = worst;
    }
    return best;
}

}  // extern "C"
