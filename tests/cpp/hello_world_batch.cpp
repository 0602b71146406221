// HelloWorld/HelloWorld.cpp:30-107 of the reference, compiled against the B200 batch engine's mirror of the
// reference API (box2dpp_b200/include/Box2D/Gpu/b2WorldBatch.h).  Differences from the reference file: the
// include line, SetContinuousPhysics(false) (the golden answers this program is compared with were frozen with continuous
// physics off; with it on — the default, like the reference's — the Python parity tests cover the same scene), and the
// output format — bit patterns instead of %4.2f so that the Python test can compare with the golden answers bit-for-bit.
//
// Modes:  hello           the reference's main(): one b2World, 60 steps
//         batch N ROWS    N pyramid worlds built through CreateBody / CreateFixture in a b2WorldBatch, with a
//                         contact listener; prints per-world event counts and the top box's state
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <Box2D/Gpu/b2WorldBatch.h>

using namespace box2d;

static unsigned bits(float f) {
    unsigned u;
    std::memcpy(&u, &f, 4);
    return u;
}

static int hello() {
    b2Vec<float, 2> gravity{{0.0f, -10.0f}};
    b2World world(gravity);
    world.SetContinuousPhysics(false);

    b2BodyDef groundBodyDef;
    groundBodyDef.position = {{0.0f, -10.0f}};
    b2Body* groundBody = world.CreateBody(&groundBodyDef);
    b2PolygonShape groundBox;
    groundBox.SetAsBox(50.0f, 10.0f);
    groundBody->CreateFixture(&groundBox, 0.0f);

    b2BodyDef bodyDef;
    bodyDef.type = b2BodyType::DYNAMIC_BODY;
    bodyDef.position = {{0.0f, 4.0f}};
    b2Body* body = world.CreateBody(&bodyDef);
    b2PolygonShape dynamicBox;
    dynamicBox.SetAsBox(1.0f, 1.0f);
    b2FixtureDef fixtureDef;
    fixtureDef.shape = &dynamicBox;
    fixtureDef.density = 1.0f;
    fixtureDef.friction = 0.3f;
    body->CreateFixture(&fixtureDef);

    float timeStep = 1.0f / 60.0f;
    int32_t velocityIterations = 6;
    int32_t positionIterations = 2;
    for (int32_t i = 0; i < 60; ++i) {
        world.Step(timeStep, velocityIterations, positionIterations);
        b2Vec<float, 2> position = body->GetPosition();
        float angle = body->GetAngle();
        b2Vec<float, 2> v = body->GetLinearVelocity();
        std::printf("%08x %08x %08x %08x %08x %08x\n", bits(position[b2VecX]), bits(position[b2VecY]), bits(angle),
                    bits(v[b2VecX]), bits(v[b2VecY]), bits(body->GetAngularVelocity()));
    }
    return world.GetBodyCount() == 2 && world.GetProxyCount() == 2 ? 0 : 3;
}

struct Counter : public b2ContactListener {
    int begin[64] = {0}, end[64] = {0}, pre[64] = {0}, post[64] = {0};
    void BeginContact(const b2ContactEvent& e) override { begin[e.world->GetIndex()]++; }
    void EndContact(const b2ContactEvent& e) override { end[e.world->GetIndex()]++; }
    void PreSolve(const b2ContactEvent& e) override { pre[e.world->GetIndex()]++; }
    void PostSolve(const b2ContactEvent& e, const b2ContactImpulse&) override { post[e.world->GetIndex()]++; }
};

// The Testbed "Pyramid" (ground variant A), world w gets w extra rows
static int batch(int n, int rows, int steps) {
    b2WorldBatch worlds(b2Vec<float, 2>{{0.0f, -10.0f}}, n);
    Counter counter;
    worlds.SetContactListener(&counter);
    std::vector<b2Body*> top(n);
    for (int w = 0; w < n; ++w) {
        b2World& world = worlds[w];
        world.SetContinuousPhysics(false);
        b2BodyDef gd;
        b2Body* ground = world.CreateBody(&gd);
        b2PolygonShape gs;
        gs.SetAsBox(50.0f, 10.0f, b2Vec<float, 2>{{0.0f, -10.0f}}, 0.0f);
        ground->CreateFixture(&gs, 0.0f);
        b2PolygonShape box;
        box.SetAsBox(0.5f, 0.5f);
        const int count = rows + w;
        b2Vec<float, 2> x{{-7.0f, 0.75f}};
        for (int i = 0; i < count; ++i) {
            b2Vec<float, 2> y = x;
            for (int j = i; j < count; ++j) {
                b2BodyDef bd;
                bd.type = b2BodyType::DYNAMIC_BODY;
                bd.position = y;
                b2Body* body = world.CreateBody(&bd);
                body->CreateFixture(&box, 5.0f);
                top[w] = body;
                y[0] += 1.125f;
            }
            x[0] += 0.5625f;
            x[1] += 1.25f;
        }
    }
    for (int s = 0; s < steps; ++s) worlds.Step(1.0f / 60.0f, 8, 3);
    if (worlds.GetStatus() != 0) return 4;
    for (int w = 0; w < n; ++w) {
        b2Vec<float, 2> p = top[w]->GetPosition();
        std::printf("%d %d %d %d %d %d %d %08x %08x %08x %d\n", w, worlds[w].GetBodyCount(), worlds[w].GetContactCount(),
                    counter.begin[w], counter.end[w], counter.pre[w], counter.post[w], bits(p[0]), bits(p[1]),
                    bits(top[w]->GetAngle()), top[w]->IsAwake() ? 1 : 0);
    }
    return 0;
}

// Between-step edits through the b2Body API: a small pyramid per world, poked every few steps the way a game /
// RL loop would (impulses, velocity resets, a teleport, putting a body to sleep).  tests/test_cpp_api.py replays the
// same calls on the oracle.
static int edits(int n, int steps) {
    b2WorldBatch worlds(b2Vec<float, 2>{{0.0f, -10.0f}}, n);
    std::vector<std::vector<b2Body*>> bodies(n);
    for (int w = 0; w < n; ++w) {
        b2World& world = worlds[w];
        world.SetContinuousPhysics(false);
        b2BodyDef gd;
        b2Body* ground = world.CreateBody(&gd);
        b2PolygonShape gs;
        gs.SetAsBox(50.0f, 10.0f, b2Vec<float, 2>{{0.0f, -10.0f}}, 0.0f);
        ground->CreateFixture(&gs, 0.0f);
        bodies[w].push_back(ground);
        b2PolygonShape box;
        box.SetAsBox(0.5f, 0.5f);
        const int count = 4;
        b2Vec<float, 2> x{{-7.0f, 0.75f}};
        for (int i = 0; i < count; ++i) {
            b2Vec<float, 2> y = x;
            for (int j = i; j < count; ++j) {
                b2BodyDef bd;
                bd.type = b2BodyType::DYNAMIC_BODY;
                bd.position = y;
                b2Body* body = world.CreateBody(&bd);
                body->CreateFixture(&box, 5.0f);
                bodies[w].push_back(body);
                y[0] += 1.125f;
            }
            x[0] += 0.5625f;
            x[1] += 1.25f;
        }
    }
    for (int s = 0; s < steps; ++s) {
        for (int w = 0; w < n; ++w) {
            std::vector<b2Body*>& b = bodies[w];
            const int nb = (int)b.size();
            b2Body* target = b[1 + (s + w) % (nb - 1)];
            switch ((s + 2 * w) % 7) {
                case 0: target->ApplyLinearImpulse(b2Vec<float, 2>{{3.0f + w, 12.0f}}, target->GetWorldCenter(), true); break;
                case 1: target->ApplyForce(b2Vec<float, 2>{{-150.0f, 400.0f}}, b2Vec<float, 2>{{0.25f * s, 2.0f}}, (s & 1) != 0); break;
                case 2: target->SetAngularVelocity(0.5f * w - 1.0f); break;
                case 3: if (s % 21 == 3) target->SetTransform(b2Vec<float, 2>{{-2.0f + w, 6.0f}}, 0.3f); break;
                case 4: target->ApplyTorque(25.0f, true); target->ApplyForceToCenter(b2Vec<float, 2>{{40.0f, 0.0f}}, true); break;
                case 5: if (s % 35 == 5) target->SetAwake(false); else target->ApplyAngularImpulse(-1.5f, true); break;
                case 6: target->SetLinearVelocity(b2Vec<float, 2>{{0.0f, s % 14 == 6 ? 4.0f : 0.0f}}); break;
            }
        }
        worlds.Step(1.0f / 60.0f, 8, 3);
    }
    if (worlds.GetStatus() != 0) return 4;
    for (int w = 0; w < n; ++w)
        for (size_t i = 0; i < bodies[w].size(); ++i) {
            b2Body* b = bodies[w][i];
            b2Vec<float, 2> p = b->GetPosition(), v = b->GetLinearVelocity();
            std::printf("%d %zu %08x %08x %08x %08x %08x %08x %d\n", w, i, bits(p[0]), bits(p[1]), bits(b->GetAngle()), bits(v[0]),
                        bits(v[1]), bits(b->GetAngularVelocity()), b->IsAwake() ? 1 : 0);
        }
    return 0;
}

// Joints through the reference's API: a cart on a rail (b2PrismaticJointDef, limits + motor) carrying a pole
// (b2RevoluteJointDef) and a bob on a b2DistanceJointDef, built with the defs' Initialize() helpers.
static int joints(int steps) {
    b2World world(b2Vec<float, 2>{{0.0f, -10.0f}});
    b2BodyDef gd;
    b2Body* ground = world.CreateBody(&gd);
    b2EdgeShape ge;
    ge.Set(b2Vec<float, 2>{{-40.0f, 0.0f}}, b2Vec<float, 2>{{40.0f, 0.0f}});
    ground->CreateFixture(&ge, 0.0f);
    b2BodyDef bd;
    bd.type = b2BodyType::DYNAMIC_BODY;
    bd.position = b2Vec<float, 2>{{0.0f, 3.0f}};
    b2Body* cart = world.CreateBody(&bd);
    b2PolygonShape cs;
    cs.SetAsBox(0.6f, 0.25f);
    cart->CreateFixture(&cs, 2.0f);
    b2PrismaticJointDef pj;
    pj.Initialize(ground, cart, b2Vec<float, 2>{{0.0f, 3.0f}}, b2Vec<float, 2>{{1.0f, 0.0f}});
    pj.lowerTranslation = -3.0f; pj.upperTranslation = 3.0f; pj.enableLimit = true;
    pj.maxMotorForce = 60.0f; pj.motorSpeed = 2.5f; pj.enableMotor = true;
    b2Joint* slide = world.CreateJoint(&pj);
    if (!slide) return 5;
    bd.position = b2Vec<float, 2>{{0.0f, 4.0f}};
    bd.angle = 0.2f;
    b2Body* pole = world.CreateBody(&bd);
    b2PolygonShape ps;
    ps.SetAsBox(0.08f, 1.0f);
    pole->CreateFixture(&ps, 1.0f);
    b2RevoluteJointDef rj;
    rj.Initialize(cart, pole, b2Vec<float, 2>{{0.0f, 3.0f}});
    if (!world.CreateJoint(&rj)) return 5;
    bd.position = b2Vec<float, 2>{{8.0f, 6.0f}};
    bd.angle = 0.0f;
    b2Body* bob = world.CreateBody(&bd);
    b2CircleShape bs;
    bs.SetRadius(0.35f);
    bob->CreateFixture(&bs, 2.0f);
    b2DistanceJointDef dj;
    dj.Initialize(ground, bob, b2Vec<float, 2>{{6.0f, 8.0f}}, b2Vec<float, 2>{{8.0f, 6.0f}});
    dj.frequencyHz = 3.0f; dj.dampingRatio = 0.2f;
    if (!world.CreateJoint(&dj)) return 5;
    // a flywheel on a ground pivot geared to the cart's slide (b2GearJointDef), and a crate held by a mouse joint
    bd.position = b2Vec<float, 2>{{-6.0f, 5.0f}};
    b2Body* wheel = world.CreateBody(&bd);
    b2CircleShape ws;
    ws.SetRadius(0.5f);
    wheel->CreateFixture(&ws, 3.0f);
    b2RevoluteJointDef wj;
    wj.Initialize(ground, wheel, b2Vec<float, 2>{{-6.0f, 5.0f}});
    b2Joint* wheelJoint = world.CreateJoint(&wj);
    if (!wheelJoint) return 5;
    b2GearJointDef gj;
    gj.bodyA = wheel; gj.bodyB = cart;
    gj.joint1 = wheelJoint; gj.joint2 = slide;
    gj.ratio = -2.0f;
    if (!world.CreateJoint(&gj)) return 5;
    bd.position = b2Vec<float, 2>{{12.0f, 4.0f}};
    b2Body* crate = world.CreateBody(&bd);
    b2PolygonShape crs;
    crs.SetAsBox(0.4f, 0.4f);
    crate->CreateFixture(&crs, 1.0f);
    b2MouseJointDef mj;
    mj.bodyA = ground; mj.bodyB = crate;
    mj.target = b2Vec<float, 2>{{12.2f, 4.3f}};
    mj.maxForce = 500.0f;
    if (!world.CreateJoint(&mj)) return 5;
    for (int s = 0; s < steps; ++s) world.Step(1.0f / 60.0f, 8, 3);
    b2Body* all[5] = {cart, pole, bob, wheel, crate};
    for (b2Body* b : all) {
        b2Vec<float, 2> p = b->GetPosition();
        std::printf("%08x %08x %08x\n", bits(p[0]), bits(p[1]), bits(b->GetAngle()));
    }
    return 0;
}

// Bodies created and destroyed on a running world through the reference's API (a game server spawning and despawning):
// b2World::CreateBody / b2Body::CreateFixture after Steps, b2World::DestroyBody, and again CreateBody (proxy ids reused).
static int spawn(int steps) {
    b2WorldBatch worlds(b2Vec<float, 2>{{0.0f, -10.0f}}, 2);
    worlds.SetSpawnHeadroom(8, 8, 8, 200, 0);
    std::vector<std::vector<b2Body*>> bodies(2);
    for (int w = 0; w < 2; ++w) {
        b2BodyDef gd;
        b2Body* ground = worlds[w].CreateBody(&gd);
        b2PolygonShape gs;
        gs.SetAsBox(20.0f, 1.0f, b2Vec<float, 2>{{0.0f, -1.0f}}, 0.0f);
        ground->CreateFixture(&gs, 0.0f);
        bodies[w].push_back(ground);
        b2BodyDef bd;
        bd.type = b2BodyType::DYNAMIC_BODY;
        bd.position = b2Vec<float, 2>{{0.0f, 2.0f + w}};
        b2Body* box = worlds[w].CreateBody(&bd);
        b2PolygonShape bs;
        bs.SetAsBox(0.5f, 0.5f);
        box->CreateFixture(&bs, 1.0f);
        bodies[w].push_back(box);
    }
    for (int s = 0; s < steps; ++s) {
        for (int w = 0; w < 2; ++w) {
            if (s % 20 == 10 && bodies[w].size() < 8) {  // spawn: alternately a box and a ball, above the pile
                b2BodyDef bd;
                bd.type = b2BodyType::DYNAMIC_BODY;
                bd.position = b2Vec<float, 2>{{0.3f * (s / 20) - 0.4f * w, 5.0f}};
                bd.angle = 0.1f * (s / 20);
                b2Body* nb = worlds[w].CreateBody(&bd);
                if ((s / 20) % 2 == 0) {
                    b2PolygonShape bs;
                    bs.SetAsBox(0.4f, 0.3f);
                    b2FixtureDef fd;
                    fd.shape = &bs; fd.density = 2.0f; fd.friction = 0.5f;
                    nb->CreateFixture(&fd);
                } else {
                    b2CircleShape cs;
                    cs.SetRadius(0.35f);
                    nb->CreateFixture(&cs, 1.5f);
                }
                nb->ApplyLinearImpulse(b2Vec<float, 2>{{0.5f, 0.0f}}, b2Vec<float, 2>{{bd.position[0], 5.1f}}, true);
                bodies[w].push_back(nb);
            }
            if (s == 55 + 20 * w) {  // despawn the first box; its proxy id goes to the next spawned body
                worlds[w].DestroyBody(bodies[w][1]);
                bodies[w][1] = nullptr;
            }
        }
        worlds.Step(1.0f / 60.0f, 8, 3);
    }
    if (worlds.GetStatus() != 0) return 4;
    for (int w = 0; w < 2; ++w)
        for (size_t i = 0; i < bodies[w].size(); ++i) {
            b2Body* b = bodies[w][i];
            if (!b) { std::printf("%d %zu dead\n", w, i); continue; }
            b2Vec<float, 2> p = b->GetPosition();
            std::printf("%d %zu %08x %08x %08x\n", w, i, bits(p[0]), bits(p[1]), bits(b->GetAngle()));
        }
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc >= 2 && std::strcmp(argv[1], "batch") == 0)
            return batch(argc > 2 ? std::atoi(argv[2]) : 3, argc > 3 ? std::atoi(argv[3]) : 4, argc > 4 ? std::atoi(argv[4]) : 120);
        if (argc >= 2 && std::strcmp(argv[1], "spawn") == 0) return spawn(argc > 2 ? std::atoi(argv[2]) : 160);
        if (argc >= 2 && std::strcmp(argv[1], "joints") == 0) return joints(argc > 2 ? std::atoi(argv[2]) : 200);
        if (argc >= 2 && std::strcmp(argv[1], "edits") == 0)
            return edits(argc > 2 ? std::atoi(argv[2]) : 3, argc > 3 ? std::atoi(argv[3]) : 150);
        return hello();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
