// The reference's edit calls on running worlds through the C++ mirror (Box2D/Gpu/b2WorldBatch.h):
// b2World::CreateJoint / DestroyJoint, b2Body::CreateFixture on an existing body, DestroyFixture, SetType, SetActive,
// SetMassData, b2Fixture::SetFilterData.
//
// World 0 is built completely before the first Step; world 1 gets its second fixture and its joint AFTER the batch exists
// (the running-world path).  Both are the same sequence of reference calls, so their state must agree bit for bit.
#include <cstdio>
#include <cstring>

#include "Box2D/Gpu/b2WorldBatch.h"

using namespace box2d;

static unsigned bits(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

struct CountAll : b2QueryCallback {
    int seen = 0, ofGround = 0;
    b2Body* ground = nullptr;
    bool stopAfterFirst = false;
    bool ReportFixture(b2Fixture* f) override {
        ++seen;
        if (f->GetBody() == ground) ++ofGround;
        return !stopAfterFirst;
    }
};
struct Closest : b2RayCastCallback {
    b2Fixture* fixture = nullptr;
    b2Vec<float, 2> point{{0.0f, 0.0f}}, normal{{0.0f, 0.0f}};
    float fraction = 1.0f;
    float ReportFixture(b2Fixture* f, const b2Vec<float, 2>& p, const b2Vec<float, 2>& n, float fr) override {
        fixture = f; point = p; normal = n; fraction = fr;
        return fr;
    }
};

int main() {
    b2WorldBatch batch(b2Vec<float, 2>{{0.0f, -10.0f}}, 3);
    b2Body* arm[3];
    b2Body* ground[3];
    b2Fixture* extra[3] = {nullptr, nullptr, nullptr};
    b2Joint* joint[3] = {nullptr, nullptr, nullptr};
    b2PolygonShape box, plate;
    box.SetAsBox(0.5f, 0.25f);
    plate.SetAsBox(0.2f, 0.6f, b2Vec<float, 2>{{0.6f, 0.0f}}, 0.3f);
    for (int w = 0; w < 3; ++w) {
        b2World& world = batch[w];
        b2BodyDef gd;
        ground[w] = world.CreateBody(&gd);
        b2PolygonShape gbox;
        gbox.SetAsBox(20.0f, 1.0f);
        ground[w]->CreateFixture(&gbox, 0.0f);
        b2BodyDef bd;
        bd.type = b2BodyType::DYNAMIC_BODY;
        bd.position = {{1.0f, 5.0f}};
        arm[w] = world.CreateBody(&bd);
        arm[w]->CreateFixture(&box, 1.0f);
        if (w == 0) {  // everything at scene time
            extra[w] = arm[w]->CreateFixture(&plate, 2.0f);
            b2RevoluteJointDef jd;
            jd.Initialize(ground[w], arm[w], b2Vec<float, 2>{{0.0f, 5.0f}});
            joint[w] = world.CreateJoint(&jd);
        }
    }
    if (batch.GetStatus() != 0) return 2;  // builds the batch: worlds 1 and 2 are running from here on
    for (int w = 1; w < 3; ++w) {
        extra[w] = arm[w]->CreateFixture(&plate, 2.0f);  // b2Body::CreateFixture on an existing body of a running world
        b2RevoluteJointDef jd;
        jd.Initialize(ground[w], arm[w], b2Vec<float, 2>{{0.0f, 5.0f}});
        joint[w] = batch[w].CreateJoint(&jd);             // b2World::CreateJoint on a running world
        if (!extra[w] || !joint[w]) return 3;
    }
    for (int i = 0; i < 90; ++i) batch.Step(1.0f / 60.0f, 8, 3);
    for (int w = 0; w < 2; ++w) {
        const b2Vec<float, 2> p = arm[w]->GetPosition(), v = arm[w]->GetLinearVelocity();
        std::printf("%08x %08x %08x %08x %08x %08x\n", bits(p[0]), bits(p[1]), bits(arm[w]->GetAngle()), bits(v[0]), bits(v[1]),
                    bits(arm[w]->GetAngularVelocity()));
    }
    // world 1: b2Fixture::SetDensity only tells at the next ResetMassData (box 1.0 x 0.5 at density 1, plate 0.4 x 1.2 at 2 -> 4)
    const float massBefore = arm[1]->GetMass();
    extra[1]->SetDensity(4.0f);
    const float massPending = arm[1]->GetMass();
    arm[1]->ResetMassData();
    const float massAfter = arm[1]->GetMass();
    const bool densityOk = massBefore > 1.45f && massBefore < 1.47f && massPending == massBefore && massAfter > 2.41f && massAfter < 2.43f &&
                           extra[1]->GetDensity() == 4.0f;
    // world 1: a mouse joint made on the running world (b2MouseJoint's constructor reads body B's transform) drags the arm,
    // which swings about its pivot at (0, 5) with its origin 1 m away, to the top of that circle and holds it there
    b2MouseJointDef mjd;
    mjd.bodyA = ground[1];
    mjd.bodyB = arm[1];
    const b2Vec<float, 2> grab = arm[1]->GetPosition();
    mjd.target = grab;
    mjd.maxForce = 1000.0f * arm[1]->GetMass();
    b2Joint* mouse = batch[1].CreateJoint(&mjd);
    const b2Vec<float, 2> goal{{0.0f, 6.0f}};
    if (mouse) mouse->SetTarget(goal);
    arm[1]->SetAwake(true);
    for (int i = 0; i < 45; ++i) batch.Step(1.0f / 60.0f, 8, 3);
    const b2Vec<float, 2> held = arm[1]->GetPosition();
    const bool mouseOk = mouse && grab[1] < 5.5f && held[1] > 5.95f && held[0] > -0.1f && held[0] < 0.1f;
    // world 2: the joint goes, the arm drops onto the ground; then it loses its plate, turns kinematic, is switched off and on
    batch[2].DestroyJoint(joint[2]);
    for (int i = 0; i < 60; ++i) batch.Step(1.0f / 60.0f, 8, 3);
    const float yFallen = arm[2]->GetPosition()[1];
    arm[2]->DestroyFixture(extra[2]);
    b2Filter none;
    none.maskBits = 0;
    b2MassData md;
    md.mass = 3.0f; md.I = 1.0f;
    arm[2]->SetMassData(&md);
    arm[2]->SetLinearDamping(0.1f);
    arm[2]->SetGravityScale(1.0f);
    arm[2]->SetBullet(false);
    arm[2]->SetSleepingAllowed(true);
    arm[2]->SetFixedRotation(false);
    extra[1]->SetSensor(false);
    extra[1]->SetFriction(0.2f);
    extra[1]->SetFilterData(none);
    const bool gettersOk = !extra[1]->IsSensor() && extra[1]->GetFriction() == 0.2f && extra[1]->GetRestitution() == 0.0f &&
                           extra[1]->GetFilterData().maskBits == 0 && extra[1]->GetFilterData().categoryBits == 1;
    batch.Step(1.0f / 60.0f, 8, 3);
    arm[2]->SetActive(false);
    batch.Step(1.0f / 60.0f, 8, 3);
    arm[2]->SetActive(true);
    arm[2]->SetType(b2BodyType::KINEMATIC_BODY);
    arm[2]->SetLinearVelocity(b2Vec<float, 2>{{0.0f, 1.0f}});
    for (int i = 0; i < 30; ++i) batch.Step(1.0f / 60.0f, 8, 3);
    const float yLifted = arm[2]->GetPosition()[1];
    // world 0 (ground box + the two-fixture arm on its pivot): b2World::QueryAABB over everything, then cut short; a ray
    // dropped at x = 10 from y = 10 to y = -5 meets the top of the ground box (y = 1) at fraction 0.6
    CountAll all, first;
    all.ground = first.ground = ground[0];
    first.stopAfterFirst = true;
    b2AABB everything;
    everything.lowerBound = {{-25.0f, -2.0f}};
    everything.upperBound = {{25.0f, 12.0f}};
    batch[0].QueryAABB(&all, everything);
    batch[0].QueryAABB(&first, everything);
    Closest ray;
    batch[0].RayCast(&ray, b2Vec<float, 2>{{10.0f, 10.0f}}, b2Vec<float, 2>{{10.0f, -5.0f}});
    const bool queriesOk = all.seen == 3 && all.ofGround == 1 && first.seen == 1 && ray.fixture && ray.fixture->GetBody() == ground[0] &&
                           ray.fraction > 0.59999f && ray.fraction < 0.60001f && ray.normal[1] == 1.0f && ray.point[1] > 0.9999f &&
                           ray.point[1] < 1.0001f;
    if (!queriesOk) std::printf("queries: %d fixtures (%d of the ground), %d before the stop, ray fraction %g normal.y %g\n", all.seen,
                                all.ofGround, first.seen, ray.fraction, ray.normal[1]);
    std::printf("%s\n", (yFallen < 2.5f && yLifted > yFallen + 0.4f && densityOk && mouseOk && gettersOk && queriesOk && batch.GetStatus() == 0) ? "edits ok" : "edits FAILED");
    if (!mouseOk) std::printf("mouse joint: grabbed at (%g, %g), now (%g, %g)\n", grab[0], grab[1], held[0], held[1]);
    if (!densityOk) std::printf("mass %g -> %g (pending) -> %g\n", massBefore, massPending, massAfter);
    return 0;
}
