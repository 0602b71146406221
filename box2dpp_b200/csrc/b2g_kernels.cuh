// b2g_kernels.cuh — the b2World::Step pipeline as sm_100a kernels.
//
// One thread steps one world; the 32 lanes of a warp are the 32 worlds of a tile (b2g_layout.h), so
// every access `field[item]` is a coalesced 128 B / 512 B segment across the warp and the strictly
// ordered per-world lists of the reference are walked sequentially by their
// owner lane — Gauss-Seidel order, contact-list order and island order are the reference's by
// construction.
//
// Kernel -> reference function (paths under /root/reference/Box2D, vanilla twin under
// Regression/box2d-ref/Box2DRef):
//   k_collide      b2ContactManager::Collide  Dynamics/b2ContactManager.cpp:110-175
//                  b2Contact::Update          Dynamics/Contacts/b2Contact.cpp:170-256
//                  b2ContactManager::Destroy  Dynamics/b2ContactManager.cpp:42-105
//   k_islands      b2World::Solve (DFS part)  Dynamics/b2World.cpp:381-537
//   k_solve        b2Island::Solve            Dynamics/b2Island.cpp:182-383
//                  b2ContactSolver::*         Dynamics/Contacts/b2ContactSolver.cpp:33-744
//   k_sync_pairs   b2Body::SynchronizeFixtures Dynamics/b2Body.cpp:442-453, b2Fixture::Synchronize
//                  Dynamics/b2Fixture.cpp:154-177, b2DynamicTree::MoveProxy Collision/b2DynamicTree.cpp:137-181,
//                  b2BroadPhase::UpdatePairs Collision/b2BroadPhase.h:193-247,
//                  b2ContactManager::AddPair Dynamics/b2ContactManager.cpp:182-301
#pragma once

#include "b2g_collide.cuh"
#include "b2g_layout.h"
#include "../../include/b2g_scene.h"

namespace b2g {

// Column view of one field for one world: element i lives at p[i * 32].
template <class T>
struct Col {
    T* p;
    __device__ __forceinline__ T& operator[](int i) const { return p[(size_t)i * kTile]; }
};
template <class T>
__device__ __forceinline__ Col<T> col(T* base, int tile, int cap, int lane) {
    Col<T> c;
    c.p = base + ((size_t)tile * (size_t)cap) * kTile + lane;
    return c;
}

__device__ __forceinline__ Xf LoadXf(const Col<float4>& c, int i) {
    float4 v = c[i];
    Xf xf;
    xf.p = V2(v.x, v.y);
    xf.q.s = v.z;
    xf.q.c = v.w;
    return xf;
}

// ---- shape table readers -----------------------------------------------------------------------
struct DevPoly {
    const float* w;  // word 0 of the shape for this lane; word k at w[k * 32]
    int n;
    // the shape table is immutable while a step kernel runs (written at build / spawn time only): read-only path
    __device__ __forceinline__ float ld(int k) const { return __ldg(w + k); }
    __device__ __forceinline__ DevPoly(const float* base) : w(base) { n = __float_as_int(__ldg(base)); }
    __device__ __forceinline__ int count() const { return n; }
    __device__ __forceinline__ float radius() const { return ld(1 * kTile); }
    __device__ __forceinline__ Vec2 centroid() const { return V2(ld(2 * kTile), ld(3 * kTile)); }
    __device__ __forceinline__ Vec2 vert(int i) const { return V2(ld((4 + 2 * i) * kTile), ld((5 + 2 * i) * kTile)); }
    __device__ __forceinline__ Vec2 normal(int i) const {
        return V2(ld((4 + 2 * n + 2 * i) * kTile), ld((5 + 2 * n + 2 * i) * kTile));
    }
};
// The same table entry with a vertex count known at compile time: the loops of the SAT and of the clipping unroll, every
// vertex / normal load gets a constant offset, and the whole address arithmetic of the inner loops disappears.
template <int N>
struct DevPolyN {
    const float* w;
    __device__ __forceinline__ float ld(int k) const { return __ldg(w + k); }
    __device__ __forceinline__ explicit DevPolyN(const float* base) : w(base) {}
    __device__ __forceinline__ constexpr int count() const { return N; }
    __device__ __forceinline__ float radius() const { return ld(1 * kTile); }
    __device__ __forceinline__ Vec2 centroid() const { return V2(ld(2 * kTile), ld(3 * kTile)); }
    __device__ __forceinline__ Vec2 vert(int i) const { return V2(ld((4 + 2 * i) * kTile), ld((5 + 2 * i) * kTile)); }
    __device__ __forceinline__ Vec2 normal(int i) const { return V2(ld((4 + 2 * N + 2 * i) * kTile), ld((5 + 2 * N + 2 * i) * kTile)); }
};
__device__ __forceinline__ CircleData LoadCircle(const float* w) {
    CircleData c;
    c.radius = w[0];
    c.p = V2(w[1 * kTile], w[2 * kTile]);
    return c;
}
__device__ __forceinline__ EdgeData LoadEdge(const float* w) {
    EdgeData e;
    e.radius = w[0];
    e.v0 = V2(w[1 * kTile], w[2 * kTile]);
    e.v1 = V2(w[3 * kTile], w[4 * kTile]);
    e.v2 = V2(w[5 * kTile], w[6 * kTile]);
    e.v3 = V2(w[7 * kTile], w[8 * kTile]);
    int f = __float_as_int(w[9 * kTile]);
    e.has0 = (f & 1) != 0;
    e.has3 = (f & 2) != 0;
    return e;
}
// b2ChainShape::GetChildEdge — Collision/Shapes/b2ChainShape.cpp:111-140
__device__ __forceinline__ EdgeData LoadChainChild(const float* w, const Col<float2>& verts, int index) {
    EdgeData e;
    e.radius = w[0];
    int first = __float_as_int(w[1 * kTile]);
    int count = __float_as_int(w[2 * kTile]);
    int f = __float_as_int(w[7 * kTile]);
    float2 a = verts[first + index], b = verts[first + index + 1];
    e.v1 = V2(a.x, a.y);
    e.v2 = V2(b.x, b.y);
    if (index > 0) {
        float2 p = verts[first + index - 1];
        e.v0 = V2(p.x, p.y);
        e.has0 = true;
    } else {
        e.v0 = V2(w[3 * kTile], w[4 * kTile]);
        e.has0 = (f & 1) != 0;
    }
    if (index < count - 2) {
        float2 p = verts[first + index + 2];
        e.v3 = V2(p.x, p.y);
        e.has3 = true;
    } else {
        e.v3 = V2(w[5 * kTile], w[6 * kTile]);
        e.has3 = (f & 2) != 0;
    }
    return e;
}
// m_radius of any shape
__device__ __forceinline__ float ShapeRadius(const float* w, int shapeType) {
    return shapeType == B2G_SHAPE_POLYGON ? w[1 * kTile] : w[0];
}

// Shape ComputeAABB: b2PolygonShape.cpp:335-352, b2CircleShape.cpp:85-94, b2EdgeShape.cpp:118-131,
// b2ChainShape.cpp:169-185 (no radius on chain children)
__device__ __forceinline__ void ShapeAABB(const float* w, int shapeType, const Col<float2>& chainVerts, int child, Xf xf,
                                          Vec2* lo, Vec2* hi) {
    if (shapeType == B2G_SHAPE_POLYGON) {
        DevPoly p(w);
        Vec2 lower = Mul(xf, p.vert(0));
        Vec2 upper = lower;
        for (int i = 1; i < p.count(); ++i) {
            Vec2 v = Mul(xf, p.vert(i));
            lower = Min(lower, v);
            upper = Max(upper, v);
        }
        Vec2 r = V2(p.radius(), p.radius());
        *lo = lower - r;
        *hi = upper + r;
    } else if (shapeType == B2G_SHAPE_CIRCLE) {
        CircleData c = LoadCircle(w);
        Vec2 p = xf.p + Mul(xf.q, c.p);
        *lo = V2(p.x - c.radius, p.y - c.radius);
        *hi = V2(p.x + c.radius, p.y + c.radius);
    } else if (shapeType == B2G_SHAPE_EDGE) {
        EdgeData e = LoadEdge(w);
        Vec2 v1 = Mul(xf, e.v1);
        Vec2 v2 = Mul(xf, e.v2);
        Vec2 lower = Min(v1, v2);
        Vec2 upper = Max(v1, v2);
        Vec2 r = V2(e.radius, e.radius);
        *lo = lower - r;
        *hi = upper + r;
    } else {
        int first = __float_as_int(w[1 * kTile]);
        int count = __float_as_int(w[2 * kTile]);
        int i1 = child, i2 = child + 1;
        if (i2 == count) i2 = 0;
        float2 a = chainVerts[first + i1], b = chainVerts[first + i2];
        Vec2 v1 = Mul(xf, V2(a.x, a.y));
        Vec2 v2 = Mul(xf, V2(b.x, b.y));
        *lo = Min(v1, v2);
        *hi = Max(v1, v2);
    }
}

#include "b2g_distance.cuh"

__device__ __forceinline__ bool FatOverlap(float4 a, float4 b) {  // b2TestOverlap(AABB, AABB), b2Collision.h:278-291
    float d1x = b.x - a.z, d1y = b.y - a.w;
    float d2x = a.x - b.z, d2y = a.y - b.w;
    if (d1x > 0.0f || d1y > 0.0f) return false;
    if (d2x > 0.0f || d2y > 0.0f) return false;
    return true;
}

// b2Body::SetAwake(true) — Box2D/Dynamics/b2Body.h:634-643
__device__ __forceinline__ bool WakeBody(const Col<int>& bodyFlags, const Col<float4>& bodySweep, int b) {
    int f = bodyFlags[b];
    if ((f & (kBodyAwake | kBodyDestroyed)) == 0) {
        bodyFlags[b] = f | kBodyAwake;
        float4 s = bodySweep[b];
        s.w = 0.0f;  // m_sleepTime
        bodySweep[b] = s;
        return true;
    }
    return false;
}

struct EventSink {
    Col<int4> a;
    Col<float4> b;
    int* count;
    int* status;
    int cap, mask;
    __device__ __forceinline__ void push(int type, int fixA, int childA, int fixB, int childB, int n, float4 imp) const {
        if ((mask >> type & 1) == 0) return;
        int i = *count;
        if (i < cap) {
            a[i] = make_int4(type, fixA | (childA << 16), fixB | (childB << 16), n);
            b[i] = imp;
        } else {
            *status |= kStatusEventOverflow;
        }
        *count = i + 1;
    }
};

struct WorldCtx {
    int tile, lane;
    __device__ __forceinline__ size_t s() const { return (size_t)tile * kTile + lane; }  // world-scalar index
};

// =================================================================================================
// k_collide
// =================================================================================================
__device__ __forceinline__ void EvaluateContact(const Arena& A, int tile, int lane, Manifold* m, int shapeTypes,
                                                int shapeOffA, int shapeOffB, int proxyA, Xf xfA, Xf xfB) {
    const float* wA = A.shapeWords + ((size_t)tile * A.NS + shapeOffA) * kTile + lane;
    const float* wB = A.shapeWords + ((size_t)tile * A.NS + shapeOffB) * kTile + lane;
    int tA = shapeTypes & 3, tB = (shapeTypes >> 2) & 3;
    if (tA == B2G_SHAPE_POLYGON && tB == B2G_SHAPE_POLYGON) {
#ifndef B2G_NO_BOX_FAST_PATH
        // quadrilaterals against quadrilaterals (every SetAsBox pair): the same functions with the counts as constants
        if (__float_as_int(__ldg(wA)) == 4 && __float_as_int(__ldg(wB)) == 4) CollidePolygons(m, DevPolyN<4>(wA), xfA, DevPolyN<4>(wB), xfB);
        else
#endif
        CollidePolygons(m, DevPoly(wA), xfA, DevPoly(wB), xfB);
    } else if (tA == B2G_SHAPE_CIRCLE && tB == B2G_SHAPE_CIRCLE) {
        CollideCircles(m, LoadCircle(wA), xfA, LoadCircle(wB), xfB);
    } else if (tA == B2G_SHAPE_POLYGON && tB == B2G_SHAPE_CIRCLE) {
        CollidePolygonAndCircle(m, DevPoly(wA), xfA, LoadCircle(wB), xfB);
    } else {
        EdgeData e;
        if (tA == B2G_SHAPE_CHAIN) {
            int child = (col(A.proxyInfo, tile, A.NP, lane)[proxyA] >> 16) & 0xffff;
            e = LoadChainChild(wA, col(A.chainVerts, tile, A.NV, lane), child);
        } else {
            e = LoadEdge(wA);
        }
        if (tB == B2G_SHAPE_CIRCLE) CollideEdgeAndCircle(m, e, xfA, LoadCircle(wB), xfB);
        else CollideEdgeAndPolygon(m, e, xfA, DevPoly(wB), xfB);
    }
}

// Result bits of CollideOne; one byte per contact carries them from phase 1 to phase 2 of k_collide
// (point counts are 0..2: two bits each)
enum : int { kResEvaluated = 1, kResDead = 2, kResWasTouching = 4, kResTouching = 8, kResCountShift = 4, kResOldCountShift = 6 };

// One iteration of the loop in b2ContactManager::Collide for contact k, WITHOUT its side effects on other
// objects (waking bodies, listener events): fat-AABB test -> mark for destruction, else b2Contact::Update's
// manifold evaluation and impulse matching.  Returns kRes* | pointCount << 4 | old pointCount << 6.
// b2Contact::Update (b2Contact.cpp:170-256) for contact k whose a-record is a0, without the side effects on other
// objects (SetAwake, listener events), which the callers apply in list order.
__device__ __forceinline__ int UpdateOne(const Arena& A, int tile, int lane, int k, const int4 a0) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    int4 a = a0;
    const int proxyA = a.x & 0xffff;
    const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
    int flags = a.z;
    const int pointCountOld = (a.w >> 8) & 0xff;
    const bool wasTouching = (flags & kContactTouching) != 0;
    Col<int4> cK = col(A.contactK, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    int4 kk = cK[k];
    const float4 p0 = cP0[k], p1 = cP1[k];
    const uint32_t oldKey0 = (uint32_t)kk.x, oldKey1 = (uint32_t)kk.y;
    const float oldN0 = p0.z, oldT0 = p0.w, oldN1 = p1.z, oldT1 = p1.w;
    flags |= kContactEnabled;
    Manifold m;
    m.pointCount = 0;
    m.type = a.w & 0xff;
    const int2 so = col(A.contactS, tile, A.NC, lane)[k];
    bool touching;
    if (flags & kContactSensor) {
        // sensors don't generate manifolds: touching = b2TestOverlap(shapeA, indexA, shapeB, indexB, xfA, xfB)
        Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
        const int childA = (proxyInfo[proxyA] >> 16) & 0xffff, childB = (proxyInfo[(a.x >> 16) & 0xffff] >> 16) & 0xffff;
        touching = TestOverlapShapes(&A, tile, lane, so.x, (a.w >> 16) & 3, childA, LoadXf(bodyXf, bodyA), so.y, (a.w >> 18) & 3,
                                     childB, LoadXf(bodyXf, bodyB));
    } else {
        EvaluateContact(A, tile, lane, &m, (a.w >> 16) & 0xf, so.x, so.y, proxyA, LoadXf(bodyXf, bodyA), LoadXf(bodyXf, bodyB));
        touching = m.pointCount > 0;
    }
    if (m.pointCount > 0) {
        float nI = 0.0f, tI = 0.0f;
        if (pointCountOld > 0 && oldKey0 == m.key0) { nI = oldN0; tI = oldT0; }
        else if (pointCountOld > 1 && oldKey1 == m.key0) { nI = oldN1; tI = oldT1; }
        const float4 np0 = make_float4(m.lp0.x, m.lp0.y, nI, tI);
        if (__float_as_int(np0.x) != __float_as_int(p0.x) || __float_as_int(np0.y) != __float_as_int(p0.y) ||
            __float_as_int(np0.z) != __float_as_int(p0.z) || __float_as_int(np0.w) != __float_as_int(p0.w))
            cP0[k] = np0;
        if (m.pointCount > 1) {
            nI = 0.0f; tI = 0.0f;
            if (pointCountOld > 0 && oldKey0 == m.key1) { nI = oldN0; tI = oldT0; }
            else if (pointCountOld > 1 && oldKey1 == m.key1) { nI = oldN1; tI = oldT1; }
            cP1[k] = make_float4(m.lp1.x, m.lp1.y, nI, tI);
        }
        col(A.contactN, tile, A.NC, lane)[k] = make_float4(m.localNormal.x, m.localNormal.y, m.localPoint.x, m.localPoint.y);
        // points[1].id keeps its old value while the manifold has one point (the reference evaluates in place)
        const int newKey1 = m.pointCount > 1 ? (int)m.key1 : kk.y;
        if (kk.x != (int)m.key0 || kk.y != newKey1) {
            kk.x = (int)m.key0;
            kk.y = newKey1;
            cK[k] = kk;
        }
    }
    if (touching) flags |= kContactTouching; else flags &= ~kContactTouching;
    a.z = flags;
    a.w = (a.w & 0xffff0000) | (m.type & 0xff) | (m.pointCount << 8);
    if (a.z != a0.z || a.w != a0.w) cA[k] = a;
    return kResEvaluated | (wasTouching ? kResWasTouching : 0) | (touching ? kResTouching : 0) |
           (m.pointCount << kResCountShift) | (pointCountOld << kResOldCountShift);
}

// b2ContactFilter::ShouldCollide — Dynamics/b2WorldCallbacks.cpp:26-39
__device__ __forceinline__ bool FilterShouldCollide(int2 fa, int2 fb) {
    int groupA = fa.y, groupB = fb.y;
    if (groupA == groupB && groupA != 0) return groupA > 0;
    int catA = fa.x & 0xffff, maskA = (fa.x >> 16) & 0xffff;
    int catB = fb.x & 0xffff, maskB = (fb.x >> 16) & 0xffff;
    return (maskA & catB) != 0 && (catA & maskB) != 0;
}

// b2Body::ShouldCollide (b2Body.cpp:395-416): at least one dynamic body, and no joint between the two bodies that has
// collideConnected == false
__device__ __forceinline__ bool BodiesShouldCollide(const Arena& A, int tile, int lane, int bodyA, int bodyB) {
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    if ((bodyFlags[bodyA] & kBodyTypeMask) != B2G_DYNAMIC_BODY && (bodyFlags[bodyB] & kBodyTypeMask) != B2G_DYNAMIC_BODY) return false;
    if (A.hasJoints) {
        Col<int4> jI = col(A.jointI, tile, A.NJ, lane);
        const int nj = A.jointCount[(size_t)tile * kTile + lane];
        const int ab = bodyA | (bodyB << 16), ba = bodyB | (bodyA << 16);
        for (int j = 0; j < nj; ++j) {
            const int4 ji = jI[j];
            if ((ji.y == ab || ji.y == ba) && (ji.x & (1 << 8)) == 0) return false;
        }
    }
    return true;
}

// The e_filterFlag branch of b2ContactManager::Collide (b2ContactManager.cpp:124-146) for a flagged contact: false = the
// contact is to be destroyed.  (Kept out of line: contacts are flagged by a handful of between-step edits only.)
__device__ __noinline__ bool RefilterKeeps(const Arena* Ap, int tile, int lane, int4 a) {
    const Arena& A = *Ap;
    if (!BodiesShouldCollide(A, tile, lane, (a.y >> 16) & 0xffff, a.y & 0xffff)) return false;  // bodyB->ShouldCollide(bodyA)
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int2> fixtureFilter = col(A.fixtureFilter, tile, A.NF, lane);
    return FilterShouldCollide(fixtureFilter[proxyInfo[a.x & 0xffff] & 0xffff], fixtureFilter[proxyInfo[(a.x >> 16) & 0xffff] & 0xffff]);
}

__device__ __forceinline__ int CollideOne(const Arena& A, int tile, int lane, int k, const int4 a0) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    const int proxyA = a0.x & 0xffff, proxyB = (a0.x >> 16) & 0xffff;
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    if (!FatOverlap(proxyFat[proxyA], proxyFat[proxyB])) {
        int4 a = a0;
        a.z = a0.z | kContactDead;
        cA[k] = a;
        const int pointCountOld = (a0.w >> 8) & 0xff;
        return kResDead | ((a0.z & kContactTouching) ? kResWasTouching : 0) | (pointCountOld << kResOldCountShift);
    }
    return UpdateOne(A, tile, lane, k, a0);
}

__device__ __forceinline__ void PrefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// A contact flagged for filtering is looked at before anything else, awake or not (b2ContactManager.cpp:124-146).
// Returns true when it is to be destroyed: marked dead like a contact whose fat AABBs parted, `res` says so to the
// list-order pass.  Otherwise the flag is cleared and the contact goes on as usual.
__device__ __forceinline__ bool RefilterOne(const Arena& A, int tile, int lane, int k, int4& a0, int& res) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    if (!RefilterKeeps(&A, tile, lane, a0)) {
        a0.z |= kContactDead;
        cA[k] = a0;
        res = kResDead | ((a0.z & kContactTouching) ? kResWasTouching : 0) | (((a0.w >> 8) & 3) << kResOldCountShift);
        return true;
    }
    a0.z &= ~kContactFilter;
    cA[k] = a0;
    return false;
}

constexpr int kCollideWarps = 8;
constexpr int kCollideClasses = 6;
// Which manifold function b2Contact::Update runs for this contact (EvaluateContact's dispatch; 5 = sensor overlap test)
__device__ __forceinline__ int CollideClass(const int4 a) {
    if (a.z & kContactSensor) return 5;
    const int tA = (a.w >> 16) & 3, tB = (a.w >> 18) & 3;
    if (tA == B2G_SHAPE_POLYGON && tB == B2G_SHAPE_POLYGON) return 0;
    if (tA == B2G_SHAPE_CIRCLE && tB == B2G_SHAPE_CIRCLE) return 1;
    if (tA == B2G_SHAPE_POLYGON && tB == B2G_SHAPE_CIRCLE) return 2;
    return tB == B2G_SHAPE_CIRCLE ? 3 : 4;
}

// CTA = one tile of 32 worlds (lane = world), kCollideWarps warps.
// Phase 1 (all warps, contacts strided over warps): every contact that is active under the body flags the step
//   started with is evaluated independently (CollideOne).
// Phase 2 (warp 0, only lanes that need it): the reference's list-order side effects — SetAwake on touching
//   changes and destruction, contacts that only became active because an earlier contact woke a body, listener
//   events in list order, and the stable erase of destroyed contacts.
#ifndef B2G_COLLIDE_MIN_BLOCKS
#define B2G_COLLIDE_MIN_BLOCKS 4  // 64 registers, 4 CTAs per SM: measured best of 2..6 (0.97 -> 0.83 ms at 16,384 worlds)
#endif
template <bool kResShared>
__global__ void __launch_bounds__(kCollideWarps * 32, B2G_COLLIDE_MIN_BLOCKS) k_collide(Arena A) {
    __shared__ int sSummary[kTile];
    // [NC][lane] result byte of phase 1: in shared memory when the batch steps as one group; in global memory
    // (L2-resident between the phases) when world groups run concurrently, so that this kernel's CTAs fit beside the
    // resident CTAs of another group's sequential kernels (b2g_step.cu: groups)
    extern __shared__ unsigned char sResShared[];
    unsigned char* sRes;
    if (kResShared) sRes = sResShared;
    else sRes = A.collideRes + (size_t)(A.tile0 + blockIdx.x) * A.NC * kTile;
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    if (warp == 0) {
        sSummary[lane] = 0;
        A.eventCount[ws] = 0;
    }
    const int n = A.contactCount[ws];
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    unsigned char* tmp = sRes + lane;  // contact k at tmp[k * 32]
    __syncthreads();

    int summary = 0;  // 1: the list-order pass is needed, 2: island topology changed
    // The walk is a chain of dependent loads per contact (its record -> body flags, fat AABBs -> transforms, manifold):
    // the contact-indexed columns of this warp's NEXT contact are pulled towards the SM while this one is evaluated.
    Col<int4> pfK = col(A.contactK, tile, A.NC, lane);
    Col<float4> pfP0 = col(A.contactP0, tile, A.NC, lane), pfP1 = col(A.contactP1, tile, A.NC, lane);
    Col<int2> pfS = col(A.contactS, tile, A.NC, lane);
#ifndef B2G_NO_BINNED
    if (A.collideItems != nullptr) {
        // ---- type-binned evaluation (batches whose worlds mix shape types) ----
        // With lane = world, contact k of 32 different worlds is 32 different shape pairs and the warp runs up to
        // five manifold functions one after the other.  Here the tile's active contacts are first classified and
        // counting-sorted by manifold function, then evaluated 32 items of ONE function per warp; an item is
        // (contact k, lane): the evaluation only touches that world's columns, so its result is the same.
        __shared__ int sCount[kCollideClasses], sOffset[kCollideClasses + 1], sFill[kCollideClasses], sTotal;
        uint32_t* list1 = A.collideItems + (size_t)tile * 2 * A.NC * kTile;
        uint32_t* list2 = list1 + (size_t)A.NC * kTile;
        if (threadIdx.x < kCollideClasses) { sCount[threadIdx.x] = 0; sFill[threadIdx.x] = 0; }
        if (threadIdx.x == 0) sTotal = 0;
        __syncthreads();
        Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
        for (int k = warp; k < nMax; k += kCollideWarps) {
            if (k >= n) continue;
            int4 a0 = cA[k];
            const int bfA = bodyFlags[a0.y & 0xffff], bfB = bodyFlags[(a0.y >> 16) & 0xffff];
            const bool activeA = (bfA & kBodyAwake) && (bfA & kBodyTypeMask) != B2G_STATIC_BODY;
            const bool activeB = (bfB & kBodyAwake) && (bfB & kBodyTypeMask) != B2G_STATIC_BODY;
            int res = 0;
            if ((a0.z & kContactFilter) && RefilterOne(A, tile, lane, k, a0, res)) {
                summary |= 3;
            } else if (activeA || activeB) {
                if (!FatOverlap(proxyFat[a0.x & 0xffff], proxyFat[(a0.x >> 16) & 0xffff])) {
                    res = CollideOne(A, tile, lane, k, a0);  // the destroy branch
                    summary |= 3;
                } else {
                    const int cls = CollideClass(a0);
                    list1[atomicAdd(&sTotal, 1)] = (uint32_t)k | ((uint32_t)lane << 16) | ((uint32_t)cls << 24);
                    atomicAdd(&sCount[cls], 1);
                }
                if (A.eventMask) summary |= 1;
            }
            tmp[k * kTile] = (unsigned char)res;
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // class regions padded to whole warps
            int o = 0;
            for (int c = 0; c < kCollideClasses; ++c) { sOffset[c] = o; o += (sCount[c] + 31) & ~31; }
            sOffset[kCollideClasses] = o;
        }
        __syncthreads();
        const int total = sTotal, padded = sOffset[kCollideClasses];
        for (int i = threadIdx.x; i < padded; i += kCollideWarps * 32) list2[i] = 0xffffffffu;
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += kCollideWarps * 32) {
            const uint32_t item = list1[i];
            const int c = (int)(item >> 24);
            list2[sOffset[c] + atomicAdd(&sFill[c], 1)] = item;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < padded; i += kCollideWarps * 32) {
            const uint32_t item = list2[i];
            if (item == 0xffffffffu) continue;
            const int k = (int)(item & 0xffff), ln = (int)((item >> 16) & 31);
            const int res = UpdateOne(A, tile, ln, k, col(A.contactA, tile, A.NC, ln)[k]);
            const bool changed = ((res & kResWasTouching) != 0) != ((res & kResTouching) != 0);
            if (changed) atomicOr(&sSummary[ln], 3);
            sRes[k * kTile + ln] = (unsigned char)res;
        }
    } else
#endif
    for (int k = warp; k < nMax; k += kCollideWarps) {
        if (k >= n) continue;
#ifndef B2G_NO_COLLIDE_PREFETCH
        if (k + kCollideWarps < n) {
            const int kn = k + kCollideWarps;
            PrefetchL2(&cA[kn]); PrefetchL2(&pfK[kn]); PrefetchL2(&pfP0[kn]); PrefetchL2(&pfP1[kn]); PrefetchL2(&pfS[kn]);
        }
#endif
        int4 a0 = cA[k];
        const int bodies = a0.y;
        const int bfA = bodyFlags[bodies & 0xffff], bfB = bodyFlags[(bodies >> 16) & 0xffff];
        const bool activeA = (bfA & kBodyAwake) && (bfA & kBodyTypeMask) != B2G_STATIC_BODY;
        const bool activeB = (bfB & kBodyAwake) && (bfB & kBodyTypeMask) != B2G_STATIC_BODY;
        int res = 0;
        if ((a0.z & kContactFilter) && RefilterOne(A, tile, lane, k, a0, res)) {
            summary |= 3;
        } else if (activeA || activeB) {
            res = CollideOne(A, tile, lane, k, a0);
            const bool changed = ((res & kResWasTouching) != 0) != ((res & kResTouching) != 0);
            if ((res & kResDead) || changed) summary |= 3;
            if (A.eventMask) summary |= 1;
        }
        tmp[k * kTile] = (unsigned char)res;
    }
    if (summary) atomicOr(&sSummary[lane], summary);
    __syncthreads();
    if (warp != 0) return;
    summary = sSummary[lane];
    if (summary & 2) A.topoDirty[ws] = 1;
    if ((summary & 1) == 0) return;

    // ---- list-order pass (head of the world contact list = newest = highest index) ----
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    EventSink ev{col(A.eventA, tile, A.NE, lane), col(A.eventB, tile, A.NE, lane), &A.eventCount[ws], &A.status[ws],
                 A.NE, A.eventMask};
    bool anyDead = false;
    bool anyWake = false;  // a sleeping body has been woken in this pass
    for (int k = n - 1; k >= 0; --k) {
        int res = tmp[k * kTile];
        // nothing to do in list order: evaluated, alive, touching unchanged, no listener
        if ((res & kResEvaluated) && ((res & kResWasTouching) != 0) == ((res & kResTouching) != 0) && !A.eventMask) continue;
        if (res == 0 && !anyWake) continue;  // inactive at the start of the step and nobody has been woken since
        const int4 a = cA[k];
        const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
        if (res == 0) {  // skipped in phase 1: did an earlier contact of the list wake one of its bodies?
            const int bfA = bodyFlags[bodyA], bfB = bodyFlags[bodyB];
            const bool activeA = (bfA & kBodyAwake) && (bfA & kBodyTypeMask) != B2G_STATIC_BODY;
            const bool activeB = (bfB & kBodyAwake) && (bfB & kBodyTypeMask) != B2G_STATIC_BODY;
            if (!activeA && !activeB) continue;
            res = CollideOne(A, tile, lane, k, a);
            A.topoDirty[ws] = 1;
        }
        const int pointCount = (res >> kResCountShift) & 3, pointCountOld = (res >> kResOldCountShift) & 3;
        const bool wasTouching = (res & kResWasTouching) != 0, touching = (res & kResTouching) != 0;
        const int proxyA = a.x & 0xffff, proxyB = (a.x >> 16) & 0xffff;
        if (res & kResDead) {
            // b2ContactManager::Destroy + b2Contact::Destroy
            if (wasTouching) {
                int pa = proxyInfo[proxyA], pb = proxyInfo[proxyB];
                ev.push(kEventEnd, pa & 0xffff, (pa >> 16) & 0xffff, pb & 0xffff, (pb >> 16) & 0xffff, pointCountOld,
                        make_float4(0, 0, 0, 0));
            }
            if (pointCountOld > 0) {
                anyWake |= WakeBody(bodyFlags, bodySweep, bodyA);
                anyWake |= WakeBody(bodyFlags, bodySweep, bodyB);
            }
            anyDead = true;
            continue;
        }
        const bool sensor = (a.z & kContactSensor) != 0;
        if (touching != wasTouching && !sensor) {  // the wake-up sits in Update's non-sensor branch (b2Contact.cpp:225-229)
            anyWake |= WakeBody(bodyFlags, bodySweep, bodyA);
            anyWake |= WakeBody(bodyFlags, bodySweep, bodyB);
        }
        if (A.eventMask) {
            int pa = proxyInfo[proxyA], pb = proxyInfo[proxyB];
            int fA = pa & 0xffff, chA = (pa >> 16) & 0xffff, fB = pb & 0xffff, chB = (pb >> 16) & 0xffff;
            if (!wasTouching && touching) ev.push(kEventBegin, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
            if (wasTouching && !touching) ev.push(kEventEnd, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
            if (touching && !sensor) ev.push(kEventPreSolve, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
        }
    }
    if (anyDead) {  // stable erase keeps every list in the reference's order
        Col<int4> cK = col(A.contactK, tile, A.NC, lane);
        Col<int2> cS = col(A.contactS, tile, A.NC, lane);
        Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
        Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
        Col<float4> cN = col(A.contactN, tile, A.NC, lane);
        int wpos = 0;
        for (int k = 0; k < n; ++k) {
            int4 a = cA[k];
            if (a.z & kContactDead) continue;
            if (wpos != k) {
                cA[wpos] = a;
                cK[wpos] = cK[k];
                cS[wpos] = cS[k];
                cP0[wpos] = cP0[k];
                cP1[wpos] = cP1[k];
                cN[wpos] = cN[k];
            }
            ++wpos;
        }
        A.contactCount[ws] = wpos;
    }
}

// =================================================================================================
// k_islands — b2World::Solve's island construction (Dynamics/b2World.cpp:381-537)
// =================================================================================================
// The DFS of one world is a serial pointer chase, so its latency — not its throughput — sets the kernel's
// duration.  A CTA takes a group of L consecutive worlds (L = 4..32, chosen by the host so that the group's
// contact graph fits in shared memory), loads the graph with all its threads, then one thread per world builds
// the per-body contact lists (head insertion, like b2ContactManager::AddPair does: newest contact first) and
// runs the DFS out of shared memory; island arrays go straight to global memory.
//
// Islands are a function of the contact graph, the touching / enabled flags and the awake / active flags.
// Every kernel that changes one of those raises topoDirty; a world whose flag is clear keeps last step's
// islands (body order, contact order, island flags), which are exactly what b2World::Solve would rebuild.
constexpr int kIslandThreads = 128;
enum : int { kIbType = 3, kIbAwake = 4, kIbActive = 8, kIbIsland = 16, kIbWoken = 32 };  // smem body flag byte
enum : int { kIcEligible = 1, kIcIsland = 2 };                                              // smem contact flag byte

__host__ __device__ inline size_t IslandSmemPerWorld(int NB, int NC, int NJ) {
    // next[2*NC] u16, pair[NC] 2 x u16, contact flag u8, head[NB] u16, stack[NB] u16, body flag u8;
    // with joints: the same three arrays per joint (NJ * 9, rounded up to keep the u16 arrays aligned) and a joint head per body
    return (size_t)NC * 9 + (size_t)NB * 5 + (NJ > 0 ? (size_t)((NJ + 1) & ~1) * 9 + (size_t)NB * 2 : 0);
}

// One CTA per tile of 32 worlds.  In the steady state most worlds' islands are still valid, so the CTA first compacts the
// tile's DIRTY worlds (ballot) and then walks them L at a time (L = what fits shared memory): a settled batch in which a
// fifth of the worlds changed a touching flag runs one round per tile instead of a CTA per L worlds, dirty or not.
__global__ void __launch_bounds__(kIslandThreads) k_islands(Arena A, int L, int numTiles) {
    extern __shared__ __align__(16) unsigned char ismem[];
    __shared__ int sList[kTile];
    __shared__ int sCount;
    const int t = threadIdx.x;
    const int NB = A.NB, NC = A.NC;
    // shared arrays, element (index, l) at [index * L + l]
    uint16_t* sNext = reinterpret_cast<uint16_t*>(ismem);              // [2 * NC][L]
    uint16_t* sPair = sNext + (size_t)2 * NC * L;                      // [2 * NC][L]: bodyA at 2k, bodyB at 2k + 1
    uint16_t* sHead = sPair + (size_t)2 * NC * L;                      // [NB][L]
    uint16_t* sStack = sHead + (size_t)NB * L;                         // [NB][L]
    // joints (only when the batch has any): b2Body::m_jointList as head-inserted edge lists, like the contacts
    const int NJ = A.NJ, NJe = (A.NJ + 1) & ~1;
    const bool joints = A.hasJoints != 0;
    uint16_t* sJNext = sStack + (size_t)NB * L;                        // [2 * NJe][L]
    uint16_t* sJPair = sJNext + (joints ? (size_t)2 * NJe * L : 0);    // [2 * NJe][L]
    uint16_t* sJHead = sJPair + (joints ? (size_t)2 * NJe * L : 0);    // [NB][L]
    unsigned char* sCF = reinterpret_cast<unsigned char*>(sJHead + (joints ? (size_t)NB * L : 0));  // [NC][L]
    unsigned char* sBF = sCF + (size_t)NC * L;                         // [NB][L]
    unsigned char* sJF = sBF + (size_t)NB * L;                         // [NJe][L]: joint already in an island
    (void)NJ;

    if (blockIdx.x >= numTiles) return;
    const int tile = A.tile0 + blockIdx.x;
    if (t < kTile) {
        const int w = tile * kTile + t;
        int d = 0;
        if (w < A.numWorlds) {
            d = A.topoDirty[w] | A.alwaysRebuildIslands;
            if (d) A.topoDirty[w] = 0;
        }
        const unsigned dirty = __ballot_sync(0xffffffffu, d != 0);
        if (d) sList[__popc(dirty & ((1u << t) - 1u))] = t;
        if (t == 0) sCount = __popc(dirty);
    }
    __syncthreads();
    const int dirtyCount = sCount;
    for (int r0 = 0; r0 < dirtyCount; r0 += L) {
    // ---- load the contact graph and the body flags of this round's worlds ----
    const int l = t % L;                 // world within the round
    const bool mine = r0 + l < dirtyCount;
    const int lane = mine ? sList[r0 + l] : 0;
    const int w = tile * kTile + lane;
    const size_t ws = (size_t)tile * kTile + lane;
    const int nb = mine ? A.bodyCount[ws] : 0;
    const int nc = mine ? A.contactCount[ws] : 0;
    Col<int4> cA = col(A.contactA, tile, NC, lane);
    Col<int> bodyFlags = col(A.bodyFlags, tile, NB, lane);
    const int stride = kIslandThreads / L;
    for (int k = t / L; k < nc; k += stride) {
        const int4 a = cA[k];
        sPair[(size_t)(2 * k) * L + l] = (uint16_t)(a.y & 0xffff);
        sPair[(size_t)(2 * k + 1) * L + l] = (uint16_t)((a.y >> 16) & 0xffff);
        sCF[(size_t)k * L + l] = ((a.z & kContactEnabled) && (a.z & kContactTouching) && !(a.z & kContactSensor)) ? kIcEligible : 0;
    }
    for (int b = t / L; b < nb; b += stride) {
        const int f = bodyFlags[b];
        sBF[(size_t)b * L + l] = (unsigned char)((f & kBodyTypeMask) | ((f & kBodyAwake) ? kIbAwake : 0) | ((f & kBodyActive) ? kIbActive : 0));
        sHead[(size_t)b * L + l] = 0xffff;
        if (joints) sJHead[(size_t)b * L + l] = 0xffff;
    }
    const int nj = (mine && joints) ? A.jointCount[ws] : 0;
    if (joints) {
        Col<int4> jI = col(A.jointI, tile, A.NJ, lane);
        for (int j = t / L; j < nj; j += stride) {
            const int bodies = jI[j].y;
            sJPair[(size_t)(2 * j) * L + l] = (uint16_t)(bodies & 0xffff);
            sJPair[(size_t)(2 * j + 1) * L + l] = (uint16_t)((bodies >> 16) & 0xffff);
            sJF[(size_t)j * L + l] = 0;
        }
    }
    __syncthreads();

    // ---- one thread per world: contact lists, then the DFS ----
    if (t < L && mine) {
        // b2Body::m_contactList: head insertion in creation order -> newest (highest index) first
        for (int k = 0; k < nc; ++k) {
            const int bA = sPair[(size_t)(2 * k) * L + l], bB = sPair[(size_t)(2 * k + 1) * L + l];
            sNext[(size_t)(2 * k) * L + l] = sHead[(size_t)bA * L + l];
            sHead[(size_t)bA * L + l] = (uint16_t)k;
            sNext[(size_t)(2 * k + 1) * L + l] = sHead[(size_t)bB * L + l];
            sHead[(size_t)bB * L + l] = (uint16_t)k;
        }
        // b2Body::m_jointList: CreateJoint links edge A into bodyA's list, then edge B into bodyB's (b2World.cpp:232-246)
        for (int j = 0; j < nj; ++j) {
            const int bA = sJPair[(size_t)(2 * j) * L + l], bB = sJPair[(size_t)(2 * j + 1) * L + l];
            sJNext[(size_t)(2 * j) * L + l] = sJHead[(size_t)bA * L + l];
            sJHead[(size_t)bA * L + l] = (uint16_t)j;
            sJNext[(size_t)(2 * j + 1) * L + l] = sJHead[(size_t)bB * L + l];
            sJHead[(size_t)bB * L + l] = (uint16_t)j;
        }
        Col<int> iBodies = col(A.islandBodies, tile, NB + NC, lane);
        Col<int> iBodyEnd = col(A.islandBodyEnd, tile, NB, lane);
        Col<int> iContacts = col(A.islandContacts, tile, NC, lane);
        Col<int> iContactEnd = col(A.islandContactEnd, tile, NB, lane);
        Col<int> iJoints = col(A.islandJoints, tile, A.NJ, lane);
        Col<int> iJointEnd = col(A.islandJointEnd, tile, NB, lane);
        int islands = 0, nBodies = 0, nContacts = 0, nJoints = 0;
        bool wokeAny = false;
        for (int seed = nb - 1; seed >= 0; --seed) {  // m_bodyList is head-inserted: newest first
            const int sf = sBF[(size_t)seed * L + l];
            if (sf & kIbIsland) continue;
            if ((sf & kIbAwake) == 0 || (sf & kIbActive) == 0) continue;
            if ((sf & kIbType) == B2G_STATIC_BODY) continue;
            int stackCount = 0;
            sStack[(size_t)(stackCount++) * L + l] = (uint16_t)seed;
            sBF[(size_t)seed * L + l] = (unsigned char)(sf | kIbIsland);
            const int islandBody0 = nBodies;
            while (stackCount > 0) {
                const int b = sStack[(size_t)(--stackCount) * L + l];
                iBodies[nBodies++] = b;
                int bf = sBF[(size_t)b * L + l];
                if ((bf & kIbAwake) == 0) {  // b2Body::SetAwake(true)
                    bf |= kIbAwake | kIbWoken;
                    sBF[(size_t)b * L + l] = (unsigned char)bf;
                    wokeAny = true;
                }
                if ((bf & kIbType) == B2G_STATIC_BODY) continue;
                for (int k = sHead[(size_t)b * L + l]; k != 0xffff;) {
                    const int bA = sPair[(size_t)(2 * k) * L + l], bB = sPair[(size_t)(2 * k + 1) * L + l];
                    const int side = bA == b ? 0 : 1;
                    const int kn = sNext[(size_t)(2 * k + side) * L + l];
                    const int cf = sCF[(size_t)k * L + l];
                    if (cf == kIcEligible) {  // touching, enabled, not yet in an island
                        iContacts[nContacts++] = k;
                        sCF[(size_t)k * L + l] = kIcEligible | kIcIsland;
                        const int other = side == 0 ? bB : bA;
                        const int of = sBF[(size_t)other * L + l];
                        if ((of & kIbIsland) == 0) {
                            sStack[(size_t)(stackCount++) * L + l] = (uint16_t)other;
                            sBF[(size_t)other * L + l] = (unsigned char)(of | kIbIsland);
                        }
                    }
                    k = kn;
                }
                // search all joints connected to this body (b2World.cpp:497-525)
                for (int j = joints ? sJHead[(size_t)b * L + l] : 0xffff; j != 0xffff;) {
                    const int bA = sJPair[(size_t)(2 * j) * L + l], bB = sJPair[(size_t)(2 * j + 1) * L + l];
                    const int side = bA == b ? 0 : 1;
                    const int jn = sJNext[(size_t)(2 * j + side) * L + l];
                    if (sJF[(size_t)j * L + l] == 0) {
                        const int other = side == 0 ? bB : bA;
                        const int of = sBF[(size_t)other * L + l];
                        if (of & kIbActive) {  // don't simulate joints connected to inactive bodies
                            iJoints[nJoints++] = j;
                            sJF[(size_t)j * L + l] = 1;
                            if ((of & kIbIsland) == 0) {
                                sStack[(size_t)(stackCount++) * L + l] = (uint16_t)other;
                                sBF[(size_t)other * L + l] = (unsigned char)(of | kIbIsland);
                            }
                        }
                    }
                    j = jn;
                }
            }
            // statics may join the next island
            for (int i = islandBody0; i < nBodies; ++i) {
                const int b = iBodies[i];
                const int f = sBF[(size_t)b * L + l];
                if ((f & kIbType) == B2G_STATIC_BODY) sBF[(size_t)b * L + l] = (unsigned char)(f & ~kIbIsland);
            }
            iBodyEnd[islands] = nBodies;
            iContactEnd[islands] = nContacts;
            if (joints) iJointEnd[islands] = nJoints;
            ++islands;
        }
        A.islandCount[ws] = islands;
        // A body woken by this DFS is a seed candidate from the next step on (seeds must be awake, b2World.cpp:415):
        // the same graph then yields a different seed and DFS order, so the islands have to be rebuilt once more.
        if (wokeAny) A.topoDirty[w] = 1;
    }
    __syncthreads();

    // ---- write the island / awake flags back ----
    Col<float4> bodySweep = col(A.bodySweep, tile, NB, lane);
    for (int b = t / L; b < nb; b += stride) {
        const int sf = sBF[(size_t)b * L + l];
        const int f = bodyFlags[b];
        int nf = (f & ~kBodyIsland) | ((sf & kIbIsland) ? kBodyIsland : 0) | ((sf & kIbAwake) ? kBodyAwake : 0);
        if (nf != f) bodyFlags[b] = nf;
        if (sf & kIbWoken) {
            float4 s = bodySweep[b];
            s.w = 0.0f;  // m_sleepTime
            bodySweep[b] = s;
        }
    }
    __syncthreads();  // the shared arrays are reused by the next group
    }
}

constexpr int kVcStride = 9;  // float4 per b2ContactVelocityConstraint

// ---- the reference's b2ContactSolver on one constraint at a time, body state in the body arrays ----
// Shared by the generic solver kernel (k_solve) and the TOI sub-step solver (b2g_toi.cuh).
// `vc`: per-world scratch of kVcStride float4 per constraint; i = constraint slot, k = contact array index.

// b2ContactSolver ctor + InitializeVelocityConstraints for one contact (b2ContactSolver.cpp:33-238)
__device__ __forceinline__ void InitVelocityConstraint(const Arena& A, int tile, int lane, const Col<float4>& vc, int i, int k,
                                                       bool warmStarting, float dtRatio) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<int4> cK = col(A.contactK, tile, A.NC, lane);
    Col<int2> cS = col(A.contactS, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> cN = col(A.contactN, tile, A.NC, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    const int4 a = cA[k];
    const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
    const int mtype = a.w & 0xff;
    const int pointCount = (a.w >> 8) & 0xff;
    const int shapeTypes = (a.w >> 16) & 0xf;
    const int4 kk = cK[k];
    const int2 so = cS[k];
    const float friction = __int_as_float(kk.z), restitution = __int_as_float(kk.w);
    const float radiusA = ShapeRadius(A.shapeWords + ((size_t)tile * A.NS + so.x) * kTile + lane, shapeTypes & 3);
    const float radiusB = ShapeRadius(A.shapeWords + ((size_t)tile * A.NS + so.y) * kTile + lane, (shapeTypes >> 2) & 3);
    const float4 msA = bodyMass[bodyA], msB = bodyMass[bodyB];
    const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
    const float4 swA = bodySweep[bodyA], swB = bodySweep[bodyB];
    const float4 veA = bodyVel[bodyA], veB = bodyVel[bodyB];
    const Vec2 cAp = V2(swA.x, swA.y), cBp = V2(swB.x, swB.y);
    const Vec2 vA = V2(veA.x, veA.y), vB = V2(veB.x, veB.y);
    const float wA = veA.z, wB = veB.z;
    Xf xfA, xfB;
    xfA.q = MakeRot(swA.z);
    xfB.q = MakeRot(swB.z);
    xfA.p = cAp - Mul(xfA.q, V2(msA.z, msA.w));
    xfB.p = cBp - Mul(xfB.q, V2(msB.z, msB.w));
    const float4 p0 = cP0[k];
    float4 p1 = make_float4(0, 0, 0, 0);
    if (pointCount > 1) p1 = cP1[k];
    const float4 nn = cN[k];
    // b2WorldManifold::Initialize (b2Collision.cpp:24-87)
    Vec2 normal, wp0, wp1 = V2(0, 0);
    if (mtype == kManifoldCircles) {
        normal = V2(1.0f, 0.0f);
        Vec2 pointA = Mul(xfA, V2(nn.z, nn.w));
        Vec2 pointB = Mul(xfB, V2(p0.x, p0.y));
        if (DistanceSquared(pointA, pointB) > kEpsilon * kEpsilon) {
            normal = pointB - pointA;
            Normalize(normal);
        }
        Vec2 ca = pointA + radiusA * normal;
        Vec2 cb = pointB - radiusB * normal;
        wp0 = 0.5f * (ca + cb);
    } else if (mtype == kManifoldFaceA) {
        normal = Mul(xfA.q, V2(nn.x, nn.y));
        Vec2 planePoint = Mul(xfA, V2(nn.z, nn.w));
        {
            Vec2 clipPoint = Mul(xfB, V2(p0.x, p0.y));
            Vec2 ca = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 cb = clipPoint - radiusB * normal;
            wp0 = 0.5f * (ca + cb);
        }
        if (pointCount > 1) {
            Vec2 clipPoint = Mul(xfB, V2(p1.x, p1.y));
            Vec2 ca = clipPoint + (radiusA - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 cb = clipPoint - radiusB * normal;
            wp1 = 0.5f * (ca + cb);
        }
    } else {
        normal = Mul(xfB.q, V2(nn.x, nn.y));
        Vec2 planePoint = Mul(xfB, V2(nn.z, nn.w));
        {
            Vec2 clipPoint = Mul(xfA, V2(p0.x, p0.y));
            Vec2 cb = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 ca = clipPoint - radiusA * normal;
            wp0 = 0.5f * (ca + cb);
        }
        if (pointCount > 1) {
            Vec2 clipPoint = Mul(xfA, V2(p1.x, p1.y));
            Vec2 cb = clipPoint + (radiusB - Dot(clipPoint - planePoint, normal)) * normal;
            Vec2 ca = clipPoint - radiusA * normal;
            wp1 = 0.5f * (ca + cb);
        }
        normal = -normal;
    }
    const Vec2 tangent = Cross(normal, 1.0f);
    int vcPointCount = pointCount;
    float4 q2, q3, q4 = make_float4(0, 0, 0, 0), q5 = make_float4(0, 0, 0, 0);
    float4 q6 = make_float4(0, 0, 0, 0), q7 = make_float4(0, 0, 0, 0);
    Vec2 rA0, rB0, rA1 = V2(0, 0), rB1 = V2(0, 0);
    {
        rA0 = wp0 - cAp;
        rB0 = wp0 - cBp;
        float rnA = Cross(rA0, normal), rnB = Cross(rB0, normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float normalMass = kNormal > 0.0f ? Div(1.0f, kNormal) : 0.0f;
        float rtA = Cross(rA0, tangent), rtB = Cross(rB0, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        float tangentMass = kTangent > 0.0f ? Div(1.0f, kTangent) : 0.0f;
        float velocityBias = 0.0f;
        float vRel = Dot(normal, vB + Cross(wB, rB0) - vA - Cross(wA, rA0));
        if (vRel < -kVelocityThreshold) velocityBias = -restitution * vRel;
        q2 = make_float4(rA0.x, rA0.y, rB0.x, rB0.y);
        q3 = make_float4(normalMass, tangentMass, velocityBias, 0.0f);
    }
    if (pointCount > 1) {
        rA1 = wp1 - cAp;
        rB1 = wp1 - cBp;
        float rnA = Cross(rA1, normal), rnB = Cross(rB1, normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float normalMass = kNormal > 0.0f ? Div(1.0f, kNormal) : 0.0f;
        float rtA = Cross(rA1, tangent), rtB = Cross(rB1, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        float tangentMass = kTangent > 0.0f ? Div(1.0f, kTangent) : 0.0f;
        float velocityBias = 0.0f;
        float vRel = Dot(normal, vB + Cross(wB, rB1) - vA - Cross(wA, rA1));
        if (vRel < -kVelocityThreshold) velocityBias = -restitution * vRel;
        q4 = make_float4(rA1.x, rA1.y, rB1.x, rB1.y);
        q5 = make_float4(normalMass, tangentMass, velocityBias, 0.0f);
        // block solver preparation
        float rn1A = Cross(rA0, normal), rn1B = Cross(rB0, normal);
        float rn2A = Cross(rA1, normal), rn2B = Cross(rB1, normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
            // b2Mat22::GetInverse with ex = (k11, k12), ey = (k12, k22)
            float ma = k11, mb = k12, mc = k12, md = k22;
            float det = ma * md - mb * mc;
            if (det != 0.0f) det = Div(1.0f, det);
            q6 = make_float4(k11, k12, k22, det * md);        // K, normalMass.ex.x
            q7 = make_float4(-det * mc, -det * mb, det * ma, 0.0f);  // nm.ex.y, nm.ey.x, nm.ey.y
        } else {
            vcPointCount = 1;
        }
    }
    float4 imp = make_float4(0, 0, 0, 0);
    if (warmStarting) imp = make_float4(dtRatio * p0.z, dtRatio * p0.w, dtRatio * p1.z, dtRatio * p1.w);
    q3.w = __int_as_float(vcPointCount | (pointCount << 8));
    q5.w = radiusA;
    q7.w = radiusB;
    const size_t o = (size_t)i * kVcStride;
    vc[o + 0] = make_float4(normal.x, normal.y, friction, __int_as_float(a.y));
    vc[o + 1] = make_float4(mA, iA, mB, iB);
    vc[o + 2] = q2;
    vc[o + 3] = q3;
    vc[o + 4] = q4;
    vc[o + 5] = q5;
    vc[o + 6] = q6;
    vc[o + 7] = q7;
    vc[o + 8] = imp;
}

// WarmStart of one constraint (b2ContactSolver.cpp:240-278)
__device__ __forceinline__ void WarmStartConstraint(const Col<float4>& vc, const Col<float4>& bodyVel, int i) {
    const size_t o = (size_t)i * kVcStride;
    const float4 q0 = vc[o + 0], q1 = vc[o + 1], q2 = vc[o + 2], q3 = vc[o + 3], imp = vc[o + 8];
    const int bodies = __float_as_int(q0.w);
    const int bodyA = bodies & 0xffff, bodyB = (bodies >> 16) & 0xffff;
    const int pointCount = __float_as_int(q3.w) & 0xff;
    const float mA = q1.x, iA = q1.y, mB = q1.z, iB = q1.w;
    float4 veA = bodyVel[bodyA], veB = bodyVel[bodyB];
    Vec2 vA = V2(veA.x, veA.y), vB = V2(veB.x, veB.y);
    float wA = veA.z, wB = veB.z;
    const Vec2 normal = V2(q0.x, q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    {
        Vec2 P = imp.x * normal + imp.y * tangent;
        wA -= iA * Cross(V2(q2.x, q2.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(q2.z, q2.w), P);
        vB = vB + mB * P;
    }
    if (pointCount > 1) {
        const float4 q4 = vc[o + 4];
        Vec2 P = imp.z * normal + imp.w * tangent;
        wA -= iA * Cross(V2(q4.x, q4.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(q4.z, q4.w), P);
        vB = vB + mB * P;
    }
    bodyVel[bodyA] = make_float4(vA.x, vA.y, wA, 0.0f);
    bodyVel[bodyB] = make_float4(vB.x, vB.y, wB, 0.0f);
}

// SolveVelocityConstraints for one constraint (b2ContactSolver.cpp:280-597)
__device__ __forceinline__ void SolveVelocityConstraint(const Col<float4>& vc, const Col<float4>& bodyVel, int i) {
    const size_t o = (size_t)i * kVcStride;
    const float4 q0 = vc[o + 0], q1 = vc[o + 1], q2 = vc[o + 2], q3 = vc[o + 3];
    float4 imp = vc[o + 8];
    const int bodies = __float_as_int(q0.w);
    const int bodyA = bodies & 0xffff, bodyB = (bodies >> 16) & 0xffff;
    const int pointCount = __float_as_int(q3.w) & 0xff;
    const float mA = q1.x, iA = q1.y, mB = q1.z, iB = q1.w;
    float4 veA = bodyVel[bodyA], veB = bodyVel[bodyB];
    Vec2 vA = V2(veA.x, veA.y), vB = V2(veB.x, veB.y);
    float wA = veA.z, wB = veB.z;
    const Vec2 normal = V2(q0.x, q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    const float friction = q0.z;
    const Vec2 rA0 = V2(q2.x, q2.y), rB0 = V2(q2.z, q2.w);
    // tangent constraints first (tangentSpeed is 0: b2Contact::SetTangentSpeed is not exposed)
    {
        Vec2 dv = vB + Cross(wB, rB0) - vA - Cross(wA, rA0);
        float vt = Dot(dv, tangent) - 0.0f;
        float lambda = q3.y * (-vt);
        float maxFriction = friction * imp.x;
        float newImpulse = Clamp(imp.y + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - imp.y;
        imp.y = newImpulse;
        Vec2 P = lambda * tangent;
        vA = vA - mA * P;
        wA -= iA * Cross(rA0, P);
        vB = vB + mB * P;
        wB += iB * Cross(rB0, P);
    }
    if (pointCount == 1) {
        Vec2 dv = vB + Cross(wB, rB0) - vA - Cross(wA, rA0);
        float vn = Dot(dv, normal);
        float lambda = -q3.x * (vn - q3.z);
        float newImpulse = Max(imp.x + lambda, 0.0f);
        lambda = newImpulse - imp.x;
        imp.x = newImpulse;
        Vec2 P = lambda * normal;
        vA = vA - mA * P;
        wA -= iA * Cross(rA0, P);
        vB = vB + mB * P;
        wB += iB * Cross(rB0, P);
    } else {
        const float4 q4 = vc[o + 4], q5 = vc[o + 5], q6 = vc[o + 6], q7 = vc[o + 7];
        const Vec2 rA1 = V2(q4.x, q4.y), rB1 = V2(q4.z, q4.w);
        {
            Vec2 dv = vB + Cross(wB, rB1) - vA - Cross(wA, rA1);
            float vt = Dot(dv, tangent) - 0.0f;
            float lambda = q5.y * (-vt);
            float maxFriction = friction * imp.z;
            float newImpulse = Clamp(imp.w + lambda, -maxFriction, maxFriction);
            lambda = newImpulse - imp.w;
            imp.w = newImpulse;
            Vec2 P = lambda * tangent;
            vA = vA - mA * P;
            wA -= iA * Cross(rA1, P);
            vB = vB + mB * P;
            wB += iB * Cross(rB1, P);
        }
        // block solver
        Vec2 a = V2(imp.x, imp.z);
        Vec2 dv1 = vB + Cross(wB, rB0) - vA - Cross(wA, rA0);
        Vec2 dv2 = vB + Cross(wB, rB1) - vA - Cross(wA, rA1);
        float vn1 = Dot(dv1, normal);
        float vn2 = Dot(dv2, normal);
        Vec2 b;
        b.x = vn1 - q3.z;
        b.y = vn2 - q5.z;
        const float k11 = q6.x, k12 = q6.y, k22 = q6.z;
        // b -= b2Mul(K, a), K.ex = (k11, k12), K.ey = (k12, k22)
        b.x -= k11 * a.x + k12 * a.y;
        b.y -= k12 * a.x + k22 * a.y;
        Vec2 x;
        bool solved = false;
        // case 1: x = -b2Mul(normalMass, b)
        x.x = -(q6.w * b.x + q7.y * b.y);
        x.y = -(q7.x * b.x + q7.z * b.y);
        if (x.x >= 0.0f && x.y >= 0.0f) solved = true;
        if (!solved) {  // case 2
            x.x = -q3.x * b.x;
            x.y = 0.0f;
            vn2 = k12 * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) solved = true;
        }
        if (!solved) {  // case 3
            x.x = 0.0f;
            x.y = -q5.x * b.y;
            vn1 = k12 * x.y + b.x;
            if (x.y >= 0.0f && vn1 >= 0.0f) solved = true;
        }
        if (!solved) {  // case 4
            x.x = 0.0f;
            x.y = 0.0f;
            if (b.x >= 0.0f && b.y >= 0.0f) solved = true;
        }
        if (solved) {
            Vec2 d = x - a;
            Vec2 P1 = d.x * normal;
            Vec2 P2 = d.y * normal;
            vA = vA - mA * (P1 + P2);
            wA -= iA * (Cross(rA0, P1) + Cross(rA1, P2));
            vB = vB + mB * (P1 + P2);
            wB += iB * (Cross(rB0, P1) + Cross(rB1, P2));
            imp.x = x.x;
            imp.z = x.y;
        }
    }
    vc[o + 8] = imp;
    bodyVel[bodyA] = make_float4(vA.x, vA.y, wA, 0.0f);
    bodyVel[bodyB] = make_float4(vB.x, vB.y, wB, 0.0f);
}

// SolvePositionConstraints / SolveTOIPositionConstraints for one constraint (b2ContactSolver.cpp:614-744, 747-836);
// folds the constraint's separations into *minSeparationIO.
__device__ __forceinline__ void SolvePositionConstraint(const Arena& A, int tile, int lane, const Col<float4>& vc, int i, int k,
                                                        float baumgarte, bool toi, int toiBodyA, int toiBodyB,
                                                        float* minSeparationIO) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> cN = col(A.contactN, tile, A.NC, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    float minSeparation = *minSeparationIO;
    const size_t o = (size_t)i * kVcStride;
    const float4 q0 = vc[o + 0], q1 = vc[o + 1];
    const float radiusA = vc[o + 5].w, radiusB = vc[o + 7].w;
    const int counts = __float_as_int(vc[o + 3].w);
    const int pointCount = (counts >> 8) & 0xff;  // pc->pointCount is the manifold's
    const int bodies = __float_as_int(q0.w);
    const int bodyA = bodies & 0xffff, bodyB = (bodies >> 16) & 0xffff;
    float mA = q1.x, iA = q1.y, mB = q1.z, iB = q1.w;
    if (toi) {  // SolveTOIPositionConstraints: only the two TOI bodies move (b2ContactSolver.cpp:747-836)
        if (bodyA != toiBodyA && bodyA != toiBodyB) { mA = 0.0f; iA = 0.0f; }
        if (bodyB != toiBodyA && bodyB != toiBodyB) { mB = 0.0f; iB = 0.0f; }
    }
    const int mtype = cA[k].w & 0xff;
    const float4 nn = cN[k];
    const float4 p0 = cP0[k];
    const float4 msA = bodyMass[bodyA], msB = bodyMass[bodyB];
    const Vec2 localCenterA = V2(msA.z, msA.w), localCenterB = V2(msB.z, msB.w);
    float4 swA = bodySweep[bodyA], swB = bodySweep[bodyB];
    Vec2 cAp = V2(swA.x, swA.y), cBp = V2(swB.x, swB.y);
    float aA = swA.z, aB = swB.z;
    const Vec2 localNormal = V2(nn.x, nn.y), localPoint = V2(nn.z, nn.w);
    for (int j = 0; j < pointCount; ++j) {
        Vec2 lp = V2(p0.x, p0.y);
        if (j == 1) {
            float4 p1 = cP1[k];
            lp = V2(p1.x, p1.y);
        }
        Xf xfA, xfB;
        xfA.q = MakeRot(aA);
        xfB.q = MakeRot(aB);
        xfA.p = cAp - Mul(xfA.q, localCenterA);
        xfB.p = cBp - Mul(xfB.q, localCenterB);
        Vec2 normal, point;
        float separation;
        if (mtype == kManifoldCircles) {
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 pointB = Mul(xfB, lp);
            normal = pointB - pointA;
            Normalize(normal);
            point = 0.5f * (pointA + pointB);
            separation = Dot(pointB - pointA, normal) - radiusA - radiusB;
        } else if (mtype == kManifoldFaceA) {
            normal = Mul(xfA.q, localNormal);
            Vec2 planePoint = Mul(xfA, localPoint);
            Vec2 clipPoint = Mul(xfB, lp);
            separation = Dot(clipPoint - planePoint, normal) - radiusA - radiusB;
            point = clipPoint;
        } else {
            normal = Mul(xfB.q, localNormal);
            Vec2 planePoint = Mul(xfB, localPoint);
            Vec2 clipPoint = Mul(xfA, lp);
            separation = Dot(clipPoint - planePoint, normal) - radiusA - radiusB;
            point = clipPoint;
            normal = -normal;
        }
        Vec2 rA = point - cAp;
        Vec2 rB = point - cBp;
        minSeparation = Min(minSeparation, separation);
        float C = Clamp(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
        float rnA = Cross(rA, normal);
        float rnB = Cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? Div(-C, K) : 0.0f;
        Vec2 P = impulse * normal;
        cAp = cAp - mA * P;
        aA -= iA * Cross(rA, P);
        cBp = cBp + mB * P;
        aB += iB * Cross(rB, P);
    }
    bodySweep[bodyA] = make_float4(cAp.x, cAp.y, aA, swA.w);
    bodySweep[bodyB] = make_float4(cBp.x, cBp.y, aB, swB.w);
    *minSeparationIO = minSeparation;
}

// =================================================================================================
// k_solve
// =================================================================================================

#include "b2g_joints.cuh"

__global__ void __launch_bounds__(128) k_solve(Arena A, float h, int velIters, int posIters) {
    int w = A.tile0 * kTile + blockIdx.x * blockDim.x + threadIdx.x;
    int tile = w >> 5, lane = w & 31;
    if (tile >= A.numTiles) return;
    size_t ws = (size_t)tile * kTile + lane;
    const int islands = A.islandCount[ws];
    const int worldFlags = A.worldFlags[ws];
    const bool warmStarting = (worldFlags & B2G_WORLD_WARM_STARTING) != 0;
    const bool allowSleep = (worldFlags & B2G_WORLD_ALLOW_SLEEP) != 0;
    const bool clearForces = (worldFlags & B2G_WORLD_CLEAR_FORCES) != 0;
    const float dtRatio = A.invDt0[ws] * h;
    const float2 g2 = A.gravity[ws];
    const Vec2 gravity = V2(g2.x, g2.y);

    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<int4> cK = col(A.contactK, tile, A.NC, lane);
    Col<int2> cS = col(A.contactS, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> cN = col(A.contactN, tile, A.NC, lane);
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    Col<float4> bodyForce = col(A.bodyForce, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float4> bodyDamp = col(A.bodyDamp, tile, A.NB, lane);
    Col<int> iBodies = col(A.islandBodies, tile, A.NB + A.NC, lane);
    Col<int> iBodyEnd = col(A.islandBodyEnd, tile, A.NB, lane);
    Col<int> iContacts = col(A.islandContacts, tile, A.NC, lane);
    Col<int> iContactEnd = col(A.islandContactEnd, tile, A.NB, lane);
    Col<float4> vc = col(A.vc, tile, A.NC * kVcStride, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    EventSink ev{col(A.eventA, tile, A.NE, lane), col(A.eventB, tile, A.NE, lane), &A.eventCount[ws], &A.status[ws],
                 A.NE, A.eventMask};

    const bool joints = A.hasJoints != 0 && A.jointCount[ws] > 0;
    const JointCols jc = MakeJointCols(A, tile, lane);
    Col<int> iJoints = col(A.islandJoints, tile, A.NJ, lane);
    Col<int> iJointEnd = col(A.islandJointEnd, tile, A.NB, lane);
    int b0 = 0, c0 = 0, j0 = 0;
    for (int isl = 0; isl < islands; ++isl) {
        const int b1 = iBodyEnd[isl], c1 = iContactEnd[isl];
        const int j1 = joints ? iJointEnd[isl] : 0;

        // ---- integrate velocities (b2Island.cpp:190-225); positions/velocities stay in the body arrays ----
        for (int i = b0; i < b1; ++i) {
            int b = iBodies[i];
            int f = bodyFlags[b];
            if ((f & kBodyTypeMask) == B2G_STATIC_BODY) continue;  // c0 == c already, never integrated
            float4 sw = bodySweep[b];
            bodySweep0[b] = make_float4(sw.x, sw.y, sw.z, 0.0f);
            if ((f & kBodyTypeMask) == B2G_DYNAMIC_BODY) {
                float4 vel = bodyVel[b];
                float4 frc = bodyForce[b];
                float4 ms = bodyMass[b];
                float4 dp = bodyDamp[b];
                Vec2 v = V2(vel.x, vel.y);
                float wv = vel.z;
                v = v + h * (dp.z * gravity + ms.x * V2(frc.x, frc.y));
                wv += h * ms.y * frc.z;
                v = Div(1.0f, 1.0f + h * dp.x) * v;
                wv *= Div(1.0f, 1.0f + h * dp.y);
                bodyVel[b] = make_float4(v.x, v.y, wv, 0.0f);
                if (clearForces && (frc.x != 0.0f || frc.y != 0.0f || frc.z != 0.0f))
                    bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }

        // ---- b2ContactSolver ctor + InitializeVelocityConstraints (b2ContactSolver.cpp:33-238) ----
        for (int i = c0; i < c1; ++i) InitVelocityConstraint(A, tile, lane, vc, i, iContacts[i], warmStarting, dtRatio);

        // ---- WarmStart (b2ContactSolver.cpp:240-278) ----
        if (warmStarting)
            for (int i = c0; i < c1; ++i) WarmStartConstraint(vc, bodyVel, i);

        // ---- joints: InitVelocityConstraints, after the contacts' warm start (b2Island.cpp:252-255) ----
        for (int i = j0; i < j1; ++i) JointInitVelocity(&jc, iJoints[i], h, dtRatio, warmStarting);

        // ---- SolveVelocityConstraints: joints, then contacts (b2Island.cpp:260-269, b2ContactSolver.cpp:280-597) ----
        for (int it = 0; it < velIters; ++it) {
            for (int i = j0; i < j1; ++i) JointSolveVelocity(&jc, iJoints[i], h);
            for (int i = c0; i < c1; ++i) SolveVelocityConstraint(vc, bodyVel, i);
        }

        // ---- StoreImpulses (b2ContactSolver.cpp:599-612) + Report/PostSolve (b2Island.cpp:517-540) ----
        for (int i = c0; i < c1; ++i) {
            const size_t o = (size_t)i * kVcStride;
            const int k = iContacts[i];
            const float4 imp = vc[o + 8];
            const int vcPointCount = __float_as_int(vc[o + 3].w) & 0xff;
            float4 p0 = cP0[k];
            p0.z = imp.x;
            p0.w = imp.y;
            cP0[k] = p0;
            if (vcPointCount > 1) {
                float4 p1 = cP1[k];
                p1.z = imp.z;
                p1.w = imp.w;
                cP1[k] = p1;
            }
        }

        // ---- integrate positions (b2Island.cpp:276-306) ----
        for (int i = b0; i < b1; ++i) {
            int b = iBodies[i];
            if ((bodyFlags[b] & kBodyTypeMask) == B2G_STATIC_BODY) continue;
            float4 sw = bodySweep[b];
            float4 vel = bodyVel[b];
            Vec2 c = V2(sw.x, sw.y);
            float a = sw.z;
            Vec2 v = V2(vel.x, vel.y);
            float wv = vel.z;
            Vec2 translation = h * v;
            if (Dot(translation, translation) > kMaxTranslationSquared) {
                float ratio = Div(kMaxTranslation, Length(translation));
                v = ratio * v;
            }
            float rotation = h * wv;
            if (rotation * rotation > kMaxRotationSquared) {
                float ratio = Div(kMaxRotation, Abs(rotation));
                wv *= ratio;
            }
            c = c + h * v;
            a += h * wv;
            bodySweep[b] = make_float4(c.x, c.y, a, sw.w);
            bodyVel[b] = make_float4(v.x, v.y, wv, 0.0f);
        }

        // ---- SolvePositionConstraints (b2Island.cpp:308-328, b2ContactSolver.cpp:614-744) ----
        bool positionSolved = false;
        for (int it = 0; it < posIters; ++it) {
            float minSeparation = 0.0f;
            for (int i = c0; i < c1; ++i)
                SolvePositionConstraint(A, tile, lane, vc, i, iContacts[i], kBaumgarte, false, -1, -1, &minSeparation);
            const bool contactsOkay = minSeparation >= -3.0f * kLinearSlop;
            bool jointsOkay = true;
            for (int i = j0; i < j1; ++i) {
                const bool jointOkay = JointSolvePosition(&jc, iJoints[i]);
                jointsOkay = jointsOkay && jointOkay;
            }
            if (contactsOkay && jointsOkay) {
                positionSolved = true;
                break;
            }
        }

        // ---- write back: SynchronizeTransform (b2Island.cpp:331-339, b2Body.h:835-839) ----
        for (int i = b0; i < b1; ++i) {
            int b = iBodies[i];
            if ((bodyFlags[b] & kBodyTypeMask) == B2G_STATIC_BODY) continue;
            float4 sw = bodySweep[b];
            float4 ms = bodyMass[b];
            Rot q = MakeRot(sw.z);
            Vec2 p = V2(sw.x, sw.y) - Mul(q, V2(ms.z, ms.w));
            bodyXf[b] = make_float4(p.x, p.y, q.s, q.c);
        }

        // ---- Report: PostSolve ----
        if (A.eventMask >> kEventPostSolve & 1) {
            for (int i = c0; i < c1; ++i) {
                const size_t o = (size_t)i * kVcStride;
                const int k = iContacts[i];
                const float4 imp = vc[o + 8];
                const int vcPointCount = __float_as_int(vc[o + 3].w) & 0xff;
                const int4 a = cA[k];
                int pa = proxyInfo[a.x & 0xffff], pb = proxyInfo[(a.x >> 16) & 0xffff];
                float4 e = make_float4(imp.x, vcPointCount > 1 ? imp.z : 0.0f, imp.y, vcPointCount > 1 ? imp.w : 0.0f);
                ev.push(kEventPostSolve, pa & 0xffff, (pa >> 16) & 0xffff, pb & 0xffff, (pb >> 16) & 0xffff, vcPointCount, e);
            }
        }

        // ---- sleep (b2Island.cpp:345-382) ----
        if (allowSleep) {
            float minSleepTime = kMaxFloat;
            const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
            const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
            for (int i = b0; i < b1; ++i) {
                int b = iBodies[i];
                int f = bodyFlags[b];
                if ((f & kBodyTypeMask) == B2G_STATIC_BODY) {
                    // islands are built and solved alternately: this island's DFS re-woke a static body that an
                    // earlier island of the same step had put to sleep (b2World.cpp:441)
                    if ((f & kBodyAwake) == 0) bodyFlags[b] = f | kBodyAwake;
                    continue;
                }
                float4 vel = bodyVel[b];
                float4 sw = bodySweep[b];
                if ((f & kBodyAutoSleep) == 0 || vel.z * vel.z > angTolSqr || Dot(V2(vel.x, vel.y), V2(vel.x, vel.y)) > linTolSqr) {
                    sw.w = 0.0f;
                    minSleepTime = 0.0f;
                } else {
                    sw.w += h;
                    minSleepTime = Min(minSleepTime, sw.w);
                }
                bodySweep[b] = sw;
            }
            if (minSleepTime >= kTimeToSleep && positionSolved) {
                A.topoDirty[ws] = 1;
                for (int i = b0; i < b1; ++i) {  // b2Body::SetAwake(false), b2Body.h:644-652
                    int b = iBodies[i];
                    bodyFlags[b] &= ~kBodyAwake;
                    float4 sw = bodySweep[b];
                    sw.w = 0.0f;
                    bodySweep[b] = sw;
                    bodyVel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            }
        }
        b0 = b1;
        c0 = c1;
        j0 = j1;
    }
}

// =================================================================================================
// k_sync_pairs
// =================================================================================================
// b2Contact::s_registers (b2Contact.cpp:42-76): 0 = no contact class, 1 = primary, 2 = swapped
__device__ __forceinline__ int ContactRegister(int t1, int t2) {
    if (t2 == B2G_SHAPE_CIRCLE) return 1;                                   // circle/edge/polygon/chain vs circle
    if (t2 == B2G_SHAPE_POLYGON) return t1 == B2G_SHAPE_CIRCLE ? 2 : 1;     // polygon, edge, chain vs polygon
    // t2 is edge or chain
    if (t1 == B2G_SHAPE_CIRCLE || t1 == B2G_SHAPE_POLYGON) return 2;
    return 0;
}

struct PairCtx {
    const Arena& A;
    int tile, lane;
    size_t ws;
    Col<int4> cA, cK;
    Col<int2> cS;
    Col<int> proxyInfo, fixtureInfo, fixtureShape, bodyFlags;
    Col<float2> fixtureMat;
    Col<int2> fixtureFilter;
    Col<float4> bodySweep;
    int nc;          // running contact count
    int ncAtStart;   // contacts that existed before this UpdatePairs
};

__device__ __forceinline__ PairCtx MakePairCtx(const Arena& A, int tile, int lane) {
    const size_t ws = (size_t)tile * kTile + lane;
    PairCtx c{A, tile, lane, ws,
              col(A.contactA, tile, A.NC, lane), col(A.contactK, tile, A.NC, lane), col(A.contactS, tile, A.NC, lane),
              col(A.proxyInfo, tile, A.NP, lane), col(A.fixtureInfo, tile, A.NF, lane), col(A.fixtureShape, tile, A.NF, lane),
              col(A.bodyFlags, tile, A.NB, lane), col(A.fixtureMat, tile, A.NF, lane), col(A.fixtureFilter, tile, A.NF, lane),
              col(A.bodySweep, tile, A.NB, lane), 0, 0};
    c.nc = A.contactCount[ws];
    c.ncAtStart = c.nc;
    return c;
}

// b2ContactManager::AddPair for proxies a < b (ascending tree id).  `knownNew`: the caller has already
// established that no contact between the two proxies exists (the walk of bodyB's edge list in the reference).
// (kept out of line: it runs for the few new pairs of a step and must not set the register budget of its callers)
__device__ __noinline__ void AddPair(PairCtx& c, int proxyA, int proxyB, bool knownNew) {
    int pa = c.proxyInfo[proxyA], pb = c.proxyInfo[proxyB];
    int fixtureA = pa & 0xffff, fixtureB = pb & 0xffff;
    int fiA = c.fixtureInfo[fixtureA], fiB = c.fixtureInfo[fixtureB];
    int bodyA = fiA & 0xffff, bodyB = fiB & 0xffff;
    if (bodyA == bodyB) return;
    if (!knownNew) {
        const int key1 = proxyA | (proxyB << 16), key2 = proxyB | (proxyA << 16);
        for (int k = 0; k < c.ncAtStart; ++k) {
            int x = c.cA[k].x;
            if (x == key1 || x == key2) return;
        }
    }
    int bfA = c.bodyFlags[bodyA], bfB = c.bodyFlags[bodyB];
    // b2Body::ShouldCollide (b2Body.cpp:395-416): at least one dynamic body, and no joint between the two bodies that
    // has collideConnected == false
    (void)bfA; (void)bfB;
    if (!BodiesShouldCollide(c.A, c.tile, c.lane, bodyA, bodyB)) return;
    if (!FilterShouldCollide(c.fixtureFilter[fixtureA], c.fixtureFilter[fixtureB])) return;
    int tA = (fiA >> 16) & 3, tB = (fiB >> 16) & 3;
    int reg = ContactRegister(tA, tB);
    if (reg == 0) return;
    if (reg == 2) {
        int t;
        t = proxyA; proxyA = proxyB; proxyB = t;
        t = fixtureA; fixtureA = fixtureB; fixtureB = t;
        t = bodyA; bodyA = bodyB; bodyB = t;
        t = tA; tA = tB; tB = t;
    }
    if (c.nc >= c.A.NC) {
        c.A.status[c.ws] |= kStatusContactOverflow;
        return;
    }
    float2 matA = c.fixtureMat[fixtureA], matB = c.fixtureMat[fixtureB];
    float friction = Sqrt(matA.x * matB.x);                      // b2MixFriction
    float restitution = matA.y > matB.y ? matA.y : matB.y;       // b2MixRestitution
    int k = c.nc++;
    c.A.topoDirty[c.ws] = 1;
    const bool sensor = ((fiA | fiB) >> 18 & 1) != 0;
    c.cA[k] = make_int4(proxyA | (proxyB << 16), bodyA | (bodyB << 16), kContactEnabled | (sensor ? kContactSensor : 0),
                        (tA | (tB << 2)) << 16);
    c.cK[k] = make_int4(0, 0, __float_as_int(friction), __float_as_int(restitution));
    c.cS[k] = make_int2(c.fixtureShape[fixtureA], c.fixtureShape[fixtureB]);
    if (!sensor) {  // wake up the bodies (b2ContactManager.cpp:291-295)
        WakeBody(c.bodyFlags, c.bodySweep, bodyA);
        WakeBody(c.bodyFlags, c.bodySweep, bodyB);
    }
}

// b2Fixture::Synchronize for one proxy (Dynamics/b2Fixture.cpp:154-177) + b2DynamicTree::MoveProxy's fat-AABB rule
// (Collision/b2DynamicTree.cpp:137-181) + b2BroadPhase::BufferMove.  Returns true when the proxy was re-fattened.
__device__ __forceinline__ bool SynchronizeProxy(const Arena& A, int tile, int lane, int p, int fixture, int child, int fi, int body) {
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    Col<float4> bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float2> chainVerts = col(A.chainVerts, tile, A.NV, lane);
    int shapeType = (fi >> 16) & 3;
    const float* sw = A.shapeWords + ((size_t)tile * A.NS + col(A.fixtureShape, tile, A.NF, lane)[fixture]) * kTile + lane;
    float4 s0 = bodySweep0[body];
    float4 ms = bodyMass[body];
    Xf xf1;
    xf1.q = MakeRot(s0.z);
    xf1.p = V2(s0.x, s0.y) - Mul(xf1.q, V2(ms.z, ms.w));
    Xf xf2 = LoadXf(bodyXf, body);
    Vec2 lo1, hi1, lo2, hi2;
    ShapeAABB(sw, shapeType, chainVerts, child, xf1, &lo1, &hi1);
    ShapeAABB(sw, shapeType, chainVerts, child, xf2, &lo2, &hi2);
    Vec2 lo = Min(lo1, lo2), hi = Max(hi1, hi2);
    Vec2 displacement = xf2.p - xf1.p;
    float4 fat = proxyFat[p];
    if (fat.x > fat.z) return false;  // a freed proxy slot (destroyed fixture, deactivated body): it stays out of the broad phase
    // b2AABB::Contains
    bool contains = fat.x <= lo.x && fat.y <= lo.y && hi.x <= fat.z && hi.y <= fat.w;
    if (contains) return false;
    Vec2 r = V2(kAabbExtension, kAabbExtension);
    lo = lo - r;
    hi = hi + r;
    Vec2 d = kAabbMultiplier * displacement;
    if (d.x < 0.0f) lo.x += d.x; else hi.x += d.x;
    if (d.y < 0.0f) lo.y += d.y; else hi.y += d.y;
    proxyFat[p] = make_float4(lo.x, lo.y, hi.x, hi.y);
    col(A.proxyMoved, tile, A.NP, lane)[p] = 1;  // b2BroadPhase::BufferMove
    return true;
}

// b2BroadPhase::UpdatePairs for one world by one thread: the move buffer is read from the proxyMoved flags and
// the pairs are produced directly in (proxyIdA, proxyIdB) order.  Used for the initial FindNewContacts of a
// freshly built world and as the overflow path of k_sync_pairs.
__device__ __noinline__ void UpdatePairsOrdered(const Arena* Ap, int tile, int lane) {
    const Arena& A = *Ap;
    const size_t ws = (size_t)tile * kTile + lane;
    const int np = A.proxyCount[ws];
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> movedList = col(A.movedList, tile, A.NP, lane);
    int nMoved = 0;
    for (int p = 0; p < np; ++p) {
        if (proxyMoved[p]) {
            movedList[nMoved++] = p;
            proxyMoved[p] = 0;
        }
    }
    if (nMoved == 0) return;
    PairCtx c = MakePairCtx(A, tile, lane);
    int mi = 0;  // first moved proxy with index > a
    for (int a = 0; a < np; ++a) {
        while (mi < nMoved && movedList[mi] <= a) ++mi;
        bool aMoved = mi > 0 && movedList[mi - 1] == a;
        float4 fatA = proxyFat[a];
        if (aMoved) {
            for (int b = a + 1; b < np; ++b)
                if (FatOverlap(proxyFat[b], fatA)) AddPair(c, a, b, false);
        } else {
            for (int m = mi; m < nMoved; ++m) {
                int b = movedList[m];
                if (FatOverlap(fatA, proxyFat[b])) AddPair(c, a, b, false);
            }
        }
    }
    A.contactCount[ws] = c.nc;
}

// Body -> proxies index (the reference reaches a body's proxies through m_fixtureList / m_proxies): counting sort
// of the proxies by body, one thread per world.  Runs when fixtures were created (with the initial FindNewContacts).
__global__ void __launch_bounds__(128) k_index_proxies(Arena A) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int tile = w >> 5, lane = w & 31;
    if (tile >= A.numTiles || w >= A.numWorlds) return;
    const size_t ws = (size_t)tile * kTile + lane;
    const int nb = A.bodyCount[ws], np = A.proxyCount[ws];
    Col<int> start = col(A.bodyProxyStart, tile, A.NB + 1, lane);
    Col<int> list = col(A.bodyProxyList, tile, A.NP, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    for (int b = 0; b <= nb; ++b) start[b] = 0;
    for (int p = 0; p < np; ++p) start[(fixtureInfo[proxyInfo[p] & 0xffff] & 0xffff) + 1] += 1;
    for (int b = 0; b < nb; ++b) start[b + 1] += start[b];  // start[b] = first slot of body b
    for (int p = 0; p < np; ++p) {
        const int b = fixtureInfo[proxyInfo[p] & 0xffff] & 0xffff;
        list[start[b]++] = p;  // start[b] ends up as the first slot of body b + 1
    }
    for (int b = nb; b > 0; --b) start[b] = start[b - 1];
    start[0] = 0;
}

// b2World::Step's leading `if (m_flags & e_newFixture) FindNewContacts()` for freshly built worlds.
__global__ void __launch_bounds__(128) k_find_new_contacts(Arena A) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int tile = w >> 5, lane = w & 31;
    if (tile >= A.numTiles || w >= A.numWorlds) return;
    const size_t ws = (size_t)tile * kTile + lane;
    if (A.newFixture[ws] == 0) return;  // only worlds whose e_newFixture is set (freshly built, or a body was created)
    A.newFixture[ws] = 0;
    UpdatePairsOrdered(&A, tile, lane);
}

// k_sync_pairs: CTA = one tile (lane = world), kPairWarps warps.
//   A  all warps, proxies strided: b2Body::SynchronizeFixtures -> b2Fixture::Synchronize -> b2DynamicTree::MoveProxy
//      (swept AABB, containment test against the fat AABB, re-fattening); moved proxies go to a per-world list.
//   B  all warps: every moved proxy is tested against every proxy of its world (the tree query of UpdatePairs);
//      hits become candidate pairs (min, max).
//   C  warp 0: sort + dedupe the candidates (std::sort + the duplicate skip of UpdatePairs).
//   D  all warps: one pass over the contact array marks candidates that already have a contact.
//   E  warp 0: b2ContactManager::AddPair for the remaining candidates in sorted order.
// Worlds whose move buffer or candidate list overflows the shared-memory lists take UpdatePairsOrdered instead.
constexpr int kPairWarps = 8;
#ifndef B2G_MAX_MOVED
#define B2G_MAX_MOVED 64
#endif
#ifndef B2G_MAX_CAND
#define B2G_MAX_CAND 256
#endif
constexpr int kMaxMoved = B2G_MAX_MOVED;     // per world, else overflow path
constexpr int kMaxCand = B2G_MAX_CAND;       // per world, else overflow path

__global__ void __launch_bounds__(kPairWarps * 32, 4) k_sync_pairs(Arena A, float dt) {
    __shared__ uint16_t sMoved[kMaxMoved][kTile];
    __shared__ uint32_t sCand[kMaxCand][kTile];
    __shared__ unsigned char sExists[kMaxCand][kTile];
    __shared__ int sNMoved[kTile], sNCand[kTile];
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const bool live = tile * kTile + lane < A.numWorlds;
    const int np = live ? A.proxyCount[ws] : 0;
    if (warp == 0) {
        sNMoved[lane] = 0;
        sNCand[lane] = 0;
    }
    __syncthreads();
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    Col<int> fixtureShape = col(A.fixtureShape, tile, A.NF, lane);
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    Col<float4> bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float2> chainVerts = col(A.chainVerts, tile, A.NV, lane);

    // ---- A: synchronize fixtures ----
    const bool pendingMoves = live && A.pendingMoves[ws] != 0;  // set by k_body_commands (SetTransform), cleared here
    const int npMax = __reduce_max_sync(0xffffffffu, np);
    for (int p = warp; p < npMax; p += kPairWarps) {
        if (p >= np) continue;
        int pi = proxyInfo[p];
        int fixture = pi & 0xffff, child = (pi >> 16) & 0xffff;
        int fi = fixtureInfo[fixture];
        int body = fi & 0xffff;
        int bf = bodyFlags[body];
        // a proxy may already sit in the move buffer: b2Body::SetTransform between two steps (k_body_commands)
        bool moved = pendingMoves && proxyMoved[p] != 0;
        if ((bf & kBodyIsland) != 0 && (bf & kBodyTypeMask) != B2G_STATIC_BODY)  // else: not in an island, did not move
            moved |= SynchronizeProxy(A, tile, lane, p, fixture, child, fi, body);
        if (!moved) continue;
        int slot = atomicAdd(&sNMoved[lane], 1);
        if (slot < kMaxMoved) sMoved[slot][lane] = (uint16_t)p;
    }
    __syncthreads();  // also orders the fat-AABB writes before the reads below
    if (warp == 0 && pendingMoves) A.pendingMoves[ws] = 0;
    const int nMoved = sNMoved[lane];
    if (__syncthreads_or(nMoved) == 0) {
        if (warp == 0 && live) {
            if (dt > 0.0f) A.invDt0[ws] = Div(1.0f, dt);
            A.stepCount[ws] += 1;
        }
        return;
    }
    const bool overflowMoved = nMoved > kMaxMoved;

    // ---- B: candidate pairs ----
    if (nMoved > 0 && !overflowMoved) {
        for (int j = 0; j < nMoved; ++j) {
            const int m = sMoved[j][lane];
            const float4 fatM = proxyFat[m];
            for (int q = warp; q < np; q += kPairWarps) {
                if (q == m) continue;
                if (!FatOverlap(fatM, proxyFat[q])) continue;
                int slot = atomicAdd(&sNCand[lane], 1);
                if (slot < kMaxCand) sCand[slot][lane] = q < m ? ((uint32_t)q << 16 | (uint32_t)m) : ((uint32_t)m << 16 | (uint32_t)q);
            }
        }
    }
    __syncthreads();
    int nCand = sNCand[lane];
    const bool overflow = overflowMoved || nCand > kMaxCand;

    // ---- C: sort + dedupe (insertion sort; the lists are short) ----
    if (warp == 0 && !overflow && nCand > 0) {
        for (int i = 1; i < nCand; ++i) {
            uint32_t v = sCand[i][lane];
            int j = i - 1;
            while (j >= 0 && sCand[j][lane] > v) {
                sCand[j + 1][lane] = sCand[j][lane];
                --j;
            }
            sCand[j + 1][lane] = v;
        }
        int u = 0;
        for (int i = 0; i < nCand; ++i) {
            uint32_t v = sCand[i][lane];
            if (u > 0 && sCand[u - 1][lane] == v) continue;
            sCand[u][lane] = v;
            sExists[u][lane] = 0;
            ++u;
        }
        nCand = u;
        sNCand[lane] = u;
    }
    __syncthreads();
    nCand = sNCand[lane];

    // ---- D: which candidates already have a contact ----
    const int nc = live ? A.contactCount[ws] : 0;
    if (!overflow && nCand > 0) {
        Col<int4> cA = col(A.contactA, tile, A.NC, lane);
        for (int k = warp; k < nc; k += kPairWarps) {
            const uint32_t x = (uint32_t)cA[k].x;  // proxyA | proxyB << 16, either order
            const uint32_t y = (x >> 16) | (x << 16);
            for (int c = 0; c < nCand; ++c) {
                const uint32_t v = sCand[c][lane];
                if (v == x || v == y) sExists[c][lane] = 1;
            }
        }
    }
    __syncthreads();

    // ---- E: AddPair in sorted order ----
    if (warp == 0 && live) {
        if (overflow) {
            UpdatePairsOrdered(&A, tile, lane);
        } else if (nMoved > 0) {
            for (int j = 0; j < nMoved; ++j) proxyMoved[sMoved[j][lane]] = 0;  // the move buffer is consumed
            if (nCand > 0) {
                PairCtx c = MakePairCtx(A, tile, lane);
                for (int i = 0; i < nCand; ++i) {
                    if (sExists[i][lane]) continue;
                    const uint32_t v = sCand[i][lane];
                    AddPair(c, (int)(v >> 16), (int)(v & 0xffff), true);
                }
                if (c.nc != c.ncAtStart) A.contactCount[ws] = c.nc;
            }
        }
        if (dt > 0.0f) A.invDt0[ws] = Div(1.0f, dt);
        A.stepCount[ws] += 1;
    }
}

}  // namespace b2g
