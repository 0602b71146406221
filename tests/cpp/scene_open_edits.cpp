// The reference's common construction patterns that touch a body BEFORE the first Step (b2Body setters and getters right
// after CreateBody, joint defs initialised from GetWorldCenter()).  They must not create the device batch nor freeze
// the scene: run without an argument this program never needs a GPU.  With "step" it also steps the world and prints
// the state of two worlds that must agree bit for bit: one built with setters, one with the same values in the defs.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Box2D/Gpu/b2WorldBatch.h"

using namespace box2d;

#define REQUIRE(c)                                                          \
    do {                                                                    \
        if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 2; } \
    } while (0)

static unsigned bits(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
    const bool step = argc > 1 && std::strcmp(argv[1], "step") == 0;
    b2WorldBatch batch(b2Vec<float, 2>{{0.0f, -10.0f}}, 2);
    b2Body* probe[2] = {nullptr, nullptr};
    for (int w = 0; w < 2; ++w) {
        b2World& world = batch[w];
        b2BodyDef gd;
        b2Body* ground = world.CreateBody(&gd);
        b2PolygonShape gbox;
        gbox.SetAsBox(20.0f, 1.0f);
        ground->CreateFixture(&gbox, 0.0f);

        b2BodyDef bd;
        bd.type = b2BodyType::DYNAMIC_BODY;
        bd.position = {{0.0f, 4.0f}};
        if (w == 1) {  // world 1: the values go through the definition
            bd.linearVelocity = {{1.5f, 0.25f}};
            bd.angularVelocity = -0.75f;
        }
        b2Body* body = world.CreateBody(&bd);
        if (w == 0) {  // world 0: the reference's usual spelling, setters right after CreateBody
            body->SetLinearVelocity(b2Vec<float, 2>{{1.5f, 0.25f}});
            body->SetAngularVelocity(-0.75f);
        }
        REQUIRE(!world.IsLocked());  // nothing above created the device batch
        REQUIRE(body->GetLinearVelocity()[0] == 1.5f && body->GetLinearVelocity()[1] == 0.25f);
        REQUIRE(body->GetAngularVelocity() == -0.75f);
        REQUIRE(body->IsAwake());
        // fixtures can still be added after the edits; an off-centre box moves the world centre
        b2PolygonShape box;
        box.SetAsBox(0.5f, 0.25f, b2Vec<float, 2>{{1.0f, 0.0f}}, 0.0f);
        b2FixtureDef fd;
        fd.shape = &box;
        fd.density = 2.0f;
        REQUIRE(body->CreateFixture(&fd) != nullptr);
        REQUIRE(!world.IsLocked());
        const b2Vec<float, 2> c = body->GetWorldCenter();
        REQUIRE(c[0] == 1.0f && c[1] == 4.0f);
        REQUIRE(body->GetPosition()[0] == 0.0f && body->GetPosition()[1] == 4.0f);
        // the other getters of b2Body.h:530-690 are served from the definition as well
        REQUIRE(body->GetMass() == 1.0f);  // 1.0 x 0.5 at density 2
        REQUIRE(body->IsActive() && body->IsSleepingAllowed() && !body->IsBullet() && !body->IsFixedRotation());
        const b2Vec<float, 2> wp = body->GetWorldPoint(b2Vec<float, 2>{{0.5f, -0.25f}}), wv = body->GetWorldVector(b2Vec<float, 2>{{0.0f, 2.0f}});
        REQUIRE(wp[0] == 0.5f && wp[1] == 3.75f && wv[0] == 0.0f && wv[1] == 2.0f);
        REQUIRE(!world.IsLocked());

        // a second body, and a joint whose def is initialised from the getters (b2RevoluteJointDef::Initialize)
        b2BodyDef bd2;
        bd2.type = b2BodyType::DYNAMIC_BODY;
        bd2.position = {{3.0f, 4.0f}};
        b2Body* arm = world.CreateBody(&bd2);
        b2CircleShape ball;
        ball.SetRadius(0.3f);
        arm->CreateFixture(&ball, 1.0f);
        b2RevoluteJointDef jd;
        jd.Initialize(body, arm, body->GetWorldCenter());
        REQUIRE(world.CreateJoint(&jd) != nullptr);
        REQUIRE(!world.IsLocked());

        // SetTransform while the scene is open is seen by the getters
        b2BodyDef bd3;
        bd3.type = b2BodyType::DYNAMIC_BODY;
        b2Body* moved = world.CreateBody(&bd3);
        moved->CreateFixture(&ball, 1.0f);
        moved->SetTransform(b2Vec<float, 2>{{-6.0f, 7.0f}}, 0.5f);
        REQUIRE(moved->GetPosition()[0] == -6.0f && moved->GetPosition()[1] == 7.0f && moved->GetAngle() == 0.5f);
        moved->SetAwake(false);
        REQUIRE(!moved->IsAwake());
        REQUIRE(!world.IsLocked());
        probe[w] = body;
    }
    if (!step) {
        std::printf("scene-open edits ok (no device batch created)\n");
        return 0;
    }
    for (int i = 0; i < 60; ++i) batch.Step(1.0f / 60.0f, 8, 3);
    for (int w = 0; w < 2; ++w) {
        const b2Vec<float, 2> p = probe[w]->GetPosition(), v = probe[w]->GetLinearVelocity();
        std::printf("%08x %08x %08x %08x %08x %08x\n", bits(p[0]), bits(p[1]), bits(probe[w]->GetAngle()), bits(v[0]), bits(v[1]),
                    bits(probe[w]->GetAngularVelocity()));
    }
    for (int w = 0; w < 2; ++w) {  // the same getters from the device: mass and flags as built, points through the body's transform
        REQUIRE(probe[w]->GetMass() == 1.0f);
        REQUIRE(probe[w]->IsActive() && probe[w]->IsSleepingAllowed() && !probe[w]->IsBullet() && !probe[w]->IsFixedRotation());
        const b2Vec<float, 2> lp{{0.5f, -0.25f}};
        const b2Vec<float, 2> back = probe[w]->GetLocalPoint(probe[w]->GetWorldPoint(lp));
        REQUIRE(back[0] > lp[0] - 1e-5f && back[0] < lp[0] + 1e-5f && back[1] > lp[1] - 1e-5f && back[1] < lp[1] + 1e-5f);
        const b2Vec<float, 2> c = probe[w]->GetWorldCenter(), v = probe[w]->GetLinearVelocity();
        const b2Vec<float, 2> vc = probe[w]->GetLinearVelocityFromWorldPoint(c);
        REQUIRE(vc[0] == v[0] && vc[1] == v[1]);
    }
    probe[0]->SetBullet(true);
    REQUIRE(probe[0]->IsBullet() && !probe[1]->IsBullet());
    // a joint that connects a body to itself is an error (on the running world too), not a silent nullptr
    try {
        b2RevoluteJointDef jd;
        jd.bodyA = probe[0];
        jd.bodyB = probe[0];
        batch[0].CreateJoint(&jd);
        std::fprintf(stderr, "CreateJoint of a body with itself did not throw\n");
        return 3;
    } catch (const std::exception&) {
    }
    return 0;
}
