// b2g_layout.h — the world-interleaved SoA arenas that replace the reference's
// b2BlockAllocator / b2StackAllocator (Box2D/Common/b2BlockAllocator.cpp, b2StackAllocator.cpp).
//
// Worlds are grouped in tiles of 32 (one warp).  Every per-world array of `cap` elements is stored
// as  elem[(tile * cap + item) * 32 + lane]  with lane = world % 32, so a warp whose lanes are 32
// consecutive worlds reads item i of a field as one contiguous 128 B (4 B elements) or 512 B
// (16 B elements) segment: fully coalesced and vectorised, and a whole tile of a field is one
// contiguous block (the unit a TMA bulk copy moves).
//
// Persistent per-world state follows SURVEY.md Appendix E; temporaries of the island solver
// (b2Island / b2ContactSolver stack allocations) live in the same arena and are reused every step.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#else
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct int2 { int x, y; };
#endif

namespace b2g {

constexpr int kTile = 32;  // worlds per tile = lanes per warp

// body flag bits (bodyFlags); low 2 bits hold b2BodyType (Box2D/Dynamics/b2Body.h:41-49, 408-417)
enum : int {
    kBodyTypeMask = 3,
    kBodyIsland = 1 << 2,
    kBodyAwake = 1 << 3,
    kBodyAutoSleep = 1 << 4,
    kBodyBullet = 1 << 5,
    kBodyFixedRotation = 1 << 6,
    kBodyActive = 1 << 7,
    kBodyToiIsland = 1 << 8,  // e_islandFlag as b2World::SolveTOI uses it (kBodyIsland keeps the discrete islands)
    kBodyToiMoved = 1 << 9,   // displaced by the current TOI sub-step (temporary)
    kFixtureHasDensity = 1 << 20, // fixtureInfo: m_density > 0 (b2Body::ResetMassData skips the others)
    kFixtureDestroyed = 1 << 19, // fixtureInfo: b2Body::DestroyFixture / DestroyBody removed the fixture (the slot is never reused)
    kBodyDestroyed = 1 << 10, // b2World::DestroyBody was called: the slot is a tombstone (never awake, never active)
};
// contact flag bits (Box2D/Dynamics/Contacts/b2Contact.h:155-174)
enum : int {
    kContactIsland = 1, kContactTouching = 2, kContactEnabled = 4, kContactFilter = 8, kContactDead = 16,
    kContactToi = 32,         // e_toiFlag: contactToi[k] holds a valid cached time of impact
    kContactToiIsland = 64,   // e_islandFlag inside b2World::SolveTOI
    kContactSensor = 128,     // fixtureA->IsSensor() || fixtureB->IsSensor(), fixed when the contact is created
    kContactToiCountShift = 8, kContactToiCountMask = 0xf << 8,  // m_toiCount (0 .. b2_maxSubSteps + 1)
};
// per-world status bits surfaced by b2gpu_batch_read_status
enum : int {
    kStatusContactOverflow = 1,
    kStatusEventOverflow = 2,
};
// event types in the readback buffer (b2ContactListener, Box2D/Dynamics/b2WorldCallbacks.h:74-120)
enum : int { kEventBegin = 0, kEventEnd = 1, kEventPreSolve = 2, kEventPostSolve = 3 };

// Words per shape in the shape table (see b2g_host.cpp for the writer):
//   polygon: [count, radius, centroid.xy, vertex0.xy .. , normal0.xy ..]          4 + 4*count words
//   circle : [radius, p.xy]                                                        3 words
//   edge   : [radius, v0.xy, v1.xy, v2.xy, v3.xy, has0 | has3 << 1]               10 words
//   chain  : [radius, firstVertex, count, prev.xy, next.xy, hasPrev | hasNext<<1]  8 words (+ vertices in chainVerts)

struct Arena {
    int numWorlds, numTiles;
    // capacities per world
    int NB, NF, NP, NC, NS, NV, NE, NJ;

    // ---- world scalars [tile][lane] ----
    float2* gravity;
    int* worldFlags;     // B2G_WORLD_* (include/b2g_scene.h)
    float* invDt0;       // b2World::m_inv_dt0
    int* bodyCount;
    int* proxyCount;
    int* contactCount;
    int* status;         // kStatus*
    int* eventCount;
    int* islandCount;    // last step (diagnostics)
    int* stepCount;
    int* fixtureCount;   // fixture slots in use
    int* shapeWordCount; // shape-table words in use
    int* chainVertCount;
    int* proxyFreeCount; // freed proxy slots (LIFO, like b2DynamicTree's free list)
    int* newFixture;     // b2World::e_newFixture
    int* pendingMoves;   // != 0: proxies were put into the move buffer between steps (b2Body::SetTransform)
    int* topoDirty;      // != 0: contact graph / touching / awake flags changed since the islands were built

    // ---- bodies [tile][NB][lane] ----
    float4* bodyXf;      // m_xf: p.x p.y q.s q.c
    float4* bodySweep;   // m_sweep.c.xy, m_sweep.a, m_sleepTime
    float4* bodySweep0;  // m_sweep.c0.xy, m_sweep.a0, m_sweep.alpha0
    float4* bodyVel;     // m_linearVelocity.xy, m_angularVelocity, (unused)
    float4* bodyForce;   // m_force.xy, m_torque, (unused)
    float4* bodyMass;    // m_invMass, m_invI, localCenter.xy
    float4* bodyDamp;    // m_linearDamping, m_angularDamping, m_gravityScale, m_mass (b2Body::GetMass)
    int* bodyFlags;      // type | kBody*
    int* bodyIsland;     // island id of this step (-1 = none)

    // ---- fixtures [tile][NF][lane] ----
    int* fixtureInfo;    // body | shapeType << 16 | isSensor << 18
    int* fixtureShape;   // word offset into shapeWords
    float2* fixtureMat;  // friction, restitution
    int2* fixtureFilter; // categoryBits | maskBits << 16, groupIndex
    float4* fixtureMass; // b2Fixture::GetMassData: mass, center.xy, I (mass 0 = density 0: b2Body::ResetMassData skips it)

    // ---- proxies [tile][NP][lane] (creation order == ascending tree id) ----
    float4* proxyFat;    // fat AABB lo.xy hi.xy (b2TreeNode::aabb)
    int* proxyInfo;      // fixture | child << 16
    int* proxyMoved;     // 1 = in the move buffer
    int* proxyFree;      // [tile][NP][lane] stack of freed proxy slots

    // ---- shapes ----
    float* shapeWords;   // [tile][NS][lane]
    float2* chainVerts;  // [tile][NV][lane]

    // ---- contacts [tile][NC][lane], array order = creation order, list order = reverse ----
    int4* contactA;      // proxyA | proxyB << 16, bodyA | bodyB << 16, flags,
                         // manifold type | pointCount << 8 | shapeTypeA << 16 | shapeTypeB << 18
    int4* contactK;      // key0, key1, friction bits, restitution bits
    int2* contactS;      // shape-table offsets of fixture A / fixture B (cached at creation)
    float4* contactP0;   // points[0]: localPoint.xy, normalImpulse, tangentImpulse
    float4* contactP1;   // points[1]
    float4* contactN;    // localNormal.xy, localPoint.xy

    // ---- joints [tile][NJ][lane], creation order (b2World::m_jointList is that array walked backwards) ----
    int* jointCount;     // [tile][lane]
    int* jointSeq;       // [tile][lane] joints created so far = the next joint id (the id rides in jointI.x bits 16-31)
    int4* jointI;        // type | collideConnected << 8 | enableLimit << 9 | enableMotor << 10, bodyA | bodyB << 16, m_limitState, 0
    float4* jointF0;     // localAnchorA.xy, localAnchorB.xy
    float4* jointF1;     // referenceAngle, lower, upper, maxMotorTorque|Force   (distance: length, frequencyHz, dampingRatio)
    float4* jointF2;     // motorSpeed, localXAxisA.xy
    float4* jointImp;    // m_impulse.xyz, m_motorImpulse
    float4* jointTmp;    // [tile][NJ * 5][lane] solver temporaries of the step (b2g_joints.cuh)
    int* islandJoints;   // [tile][NJ][lane] joint indices in solver order, concatenated over islands
    int* islandJointEnd; // [tile][NB][lane]

    // ---- per-step temporaries ----
    unsigned char* collideRes;  // [tile][NC][lane] k_collide: result byte of the evaluation pass
    uint32_t* collideItems;     // [tile][2][NC * 32] k_collide's type-binned work lists; nullptr = worlds of one shape type
    int* movedList;      // [tile][NP][lane] ascending indices of the proxies in the move buffer
    int* islandBodies;   // [tile][NB+NC][lane] bodies of all islands, concatenated (statics repeat)
    int* islandBodyEnd;  // [tile][NB][lane]   end offset into islandBodies per island
    int* islandContacts; // [tile][NC][lane]   contact indices in solver order, concatenated
    int* islandContactEnd; // [tile][NB][lane]
    // b2ContactVelocityConstraint, 9 x 16 B per island contact: [tile][NC*9][lane] (generic solver path only)
    float4* vc;
    // tile solver path (b2g_solve.cuh): per island contact i, records of all 32 worlds of a tile are contiguous
    float4* velRec;      // [tile][NC][7][lane]  velocity constraint record
    float4* posRec;      // [tile][NC][5][lane]  position constraint record
    float4* impRec;      // [tile][NC][lane]     accumulated impulses: normal0, tangent0, normal1, tangent1
    float* solveVel;     // [tile][NB][3][lane]  b2Velocity table (v.x, v.y, w): TMA source / sink of the velocity kernel
    float* solvePos;     // [tile][NB][3][lane]  b2Position table (c.x, c.y, a)
    int* solveCount;     // [tile][lane]         island contacts of this step (all islands)
    int* solveFlags;     // [tile][lane]         bit 0: some island contact has an e_circles manifold
    int* islandSolved;   // [tile][ceil(NB/32)][lane]  bit per island: position solver converged

    // ---- continuous collision (b2World::SolveTOI) ----
    float* toiMinAlpha;  // [tile][lane]          smallest time of impact of the first pass (k_toi -> k_toi_events)
    int* toiCand;        // [tile][NC][lane]      contacts that can ever be TOI candidates (static criterion), ascending
    int* toiCandCount;   // [tile][lane]
    int* bodyProxyStart; // [tile][NB+1][lane]    proxies of body b: bodyProxyList[bodyProxyStart[b] .. bodyProxyStart[b+1])
    int* bodyProxyList;  // [tile][NP][lane]      (b2Body::m_fixtureList -> b2Fixture::m_proxies, any order)
    float* contactToi;   // [tile][NC][lane]      b2Contact::m_toi (valid while kContactToi is set)
    int* toiBodies;      // [tile][64][lane]      bodies of the current TOI mini-island (2 * b2_maxTOIContacts)
    int* toiContacts;    // [tile][32][lane]      its contacts (b2_maxTOIContacts)
    float4* toiVc;       // [tile][32 * 9][lane]  its velocity constraints

    // ---- events [tile][NE][lane] ----
    int4* eventA;        // type, fixtureA | childA << 16, fixtureB | childB << 16, count
    float4* eventB;      // normalImpulses[2], tangentImpulses[2]
    int eventMask;       // bit per kEvent* type recorded
    int tile0;           // first tile of this launch (world groups run the step chain on separate streams)
    int hasJoints;       // desc.maxJoints > 0: the joint arrays are live (generic solver path)
    int alwaysRebuildIslands;  // validation switch: k_islands ignores topoDirty (desc.reserved bit 1)
};

#ifdef __CUDACC__
__host__ __device__ __forceinline__
#else
inline
#endif
size_t TileIndex(int tile, int cap, int item, int lane) {
    return ((size_t)tile * (size_t)cap + (size_t)item) * kTile + (size_t)lane;
}

}  // namespace b2g
