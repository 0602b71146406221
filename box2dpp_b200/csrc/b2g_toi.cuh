// b2g_toi.cuh — continuous collision of the Step path for sm_100a: b2World::SolveTOI and everything under it.
//
// Reference (fork paths under /root/reference/Box2D, vanilla twin under Regression/box2d-ref/Box2DRef; the two
// are bit-identical on this path):
//   b2World::SolveTOI                       Dynamics/b2World.cpp:568-886
//   b2Island::SolveTOI                      Dynamics/b2Island.cpp:385-515
//   b2ContactSolver::SolveTOIPositionConstraints  Dynamics/Contacts/b2ContactSolver.cpp:747-836
//   b2TimeOfImpact, b2SeparationFunction    Collision/b2TimeOfImpact.cpp:39-497
//   b2Distance, b2Simplex, b2DistanceProxy  Collision/b2Distance.cpp:28-604, b2Distance.h:103-141
//   b2Sweep::GetTransform/Advance/Normalize Common/b2Math.h:657-683, b2Body::Advance Dynamics/b2Body.h:841-849
//
// Shape of the kernel (k_toi, CTA = one tile of 32 worlds, lane = world):
//   phase 1  all warps, contacts strided over warps: the first pass of SolveTOI's outer loop.  Every sweep starts
//            at alpha0 = 0, so the times of impact of the candidate contacts (awake, at least one bullet or
//            non-dynamic body) are independent: b2TimeOfImpact per (world, contact) in parallel, the minimum
//            per world through a shared-memory atomicMin.
//   phase 2  (k_toi_events, one warp per tile) only lanes whose minimum is a real event (alpha < 1 - 10 eps): the reference's loop verbatim,
//            one world per lane — advance the two bodies, update the contact, build the mini-island in the
//            body's contact-list order, b2Island::SolveTOI, synchronise the displaced fixtures, FindNewContacts,
//            recompute the invalidated times of impact in list order (their sweeps now start at different
//            alpha0, which makes that part order dependent), repeat.
// Resting scenes take phase 1 only (a handful of GJK calls per world); events are rare and short.
#pragma once

#include "b2g_kernels.cuh"

namespace b2g {

constexpr float kToiBaumgarte = 0.75f;
constexpr int kMaxSubSteps = 8;
constexpr int kMaxToiContacts = 32;
constexpr int kToiBodyCap = 2 * kMaxToiContacts;
constexpr int kToiWarps = 8;

// ---- b2Sweep (b2Math.h:380-396, 657-683) --------------------------------------------------------
struct Sweep {
    Vec2 localCenter, c0, c;
    float a0, a, alpha0;
};
__device__ __forceinline__ Xf SweepTransform(const Sweep& s, float beta) {
    Xf xf;
    xf.p = (1.0f - beta) * s.c0 + beta * s.c;
    float angle = (1.0f - beta) * s.a0 + beta * s.a;
    xf.q = MakeRot(angle);
    xf.p = xf.p - Mul(xf.q, s.localCenter);
    return xf;
}
__device__ __forceinline__ void SweepAdvance(Sweep& s, float alpha) {
    float beta = Div(alpha - s.alpha0, 1.0f - s.alpha0);
    s.c0 = s.c0 + beta * (s.c - s.c0);
    s.a0 += beta * (s.a - s.a0);
    s.alpha0 = alpha;
}
__device__ __forceinline__ void SweepNormalize(Sweep& s) {
    float twoPi = 2.0f * kPi;
    float d = twoPi * floorf(Div(s.a0, twoPi));
    s.a0 -= d;
    s.a -= d;
}

// ---- b2SeparationFunction (b2TimeOfImpact.cpp:39-260) ------------------------------------------------
enum { kSepPoints = 0, kSepFaceA = 1, kSepFaceB = 2 };
struct SeparationFunction {
    const DistanceProxy* proxyA;
    const DistanceProxy* proxyB;
    Sweep sweepA, sweepB;
    int type;
    Vec2 localPoint, axis;

    __device__ void initialize(const SimplexCache* cache, const DistanceProxy* pA, const Sweep& sA, const DistanceProxy* pB,
                               const Sweep& sB, float t1) {
        proxyA = pA;
        proxyB = pB;
        const int count = cache->count;
        sweepA = sA;
        sweepB = sB;
        localPoint = V2(0.0f, 0.0f);
        Xf xfA = SweepTransform(sweepA, t1), xfB = SweepTransform(sweepB, t1);
        if (count == 1) {
            type = kSepPoints;
            Vec2 pointA = Mul(xfA, proxyA->vertex(cache->indexA[0]));
            Vec2 pointB = Mul(xfB, proxyB->vertex(cache->indexB[0]));
            axis = pointB - pointA;
            Normalize(axis);
        } else if (cache->indexA[0] == cache->indexA[1]) {
            type = kSepFaceB;  // two points on B and one on A
            Vec2 localPointB1 = proxyB->vertex(cache->indexB[0]);
            Vec2 localPointB2 = proxyB->vertex(cache->indexB[1]);
            axis = Cross(localPointB2 - localPointB1, 1.0f);
            Normalize(axis);
            Vec2 normal = Mul(xfB.q, axis);
            localPoint = 0.5f * (localPointB1 + localPointB2);
            Vec2 pointB = Mul(xfB, localPoint);
            Vec2 pointA = Mul(xfA, proxyA->vertex(cache->indexA[0]));
            float s = Dot(pointA - pointB, normal);
            if (s < 0.0f) axis = -axis;
        } else {
            type = kSepFaceA;  // two points on A and one or two points on B
            Vec2 localPointA1 = proxyA->vertex(cache->indexA[0]);
            Vec2 localPointA2 = proxyA->vertex(cache->indexA[1]);
            axis = Cross(localPointA2 - localPointA1, 1.0f);
            Normalize(axis);
            Vec2 normal = Mul(xfA.q, axis);
            localPoint = 0.5f * (localPointA1 + localPointA2);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 pointB = Mul(xfB, proxyB->vertex(cache->indexB[0]));
            float s = Dot(pointB - pointA, normal);
            if (s < 0.0f) axis = -axis;
        }
    }
    __device__ float findMinSeparation(int* indexA, int* indexB, float t) const {
        Xf xfA = SweepTransform(sweepA, t), xfB = SweepTransform(sweepB, t);
        if (type == kSepPoints) {
            Vec2 axisA = MulT(xfA.q, axis);
            Vec2 axisB = MulT(xfB.q, -axis);
            *indexA = proxyA->support(axisA);
            *indexB = proxyB->support(axisB);
            Vec2 pointA = Mul(xfA, proxyA->vertex(*indexA));
            Vec2 pointB = Mul(xfB, proxyB->vertex(*indexB));
            return Dot(pointB - pointA, axis);
        }
        if (type == kSepFaceA) {
            Vec2 normal = Mul(xfA.q, axis);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 axisB = MulT(xfB.q, -normal);
            *indexA = -1;
            *indexB = proxyB->support(axisB);
            Vec2 pointB = Mul(xfB, proxyB->vertex(*indexB));
            return Dot(pointB - pointA, normal);
        }
        Vec2 normal = Mul(xfB.q, axis);
        Vec2 pointB = Mul(xfB, localPoint);
        Vec2 axisA = MulT(xfA.q, -normal);
        *indexB = -1;
        *indexA = proxyA->support(axisA);
        Vec2 pointA = Mul(xfA, proxyA->vertex(*indexA));
        return Dot(pointA - pointB, normal);
    }
    __device__ float evaluate(int indexA, int indexB, float t) const {
        Xf xfA = SweepTransform(sweepA, t), xfB = SweepTransform(sweepB, t);
        if (type == kSepPoints) {
            Vec2 pointA = Mul(xfA, proxyA->vertex(indexA));
            Vec2 pointB = Mul(xfB, proxyB->vertex(indexB));
            return Dot(pointB - pointA, axis);
        }
        if (type == kSepFaceA) {
            Vec2 normal = Mul(xfA.q, axis);
            Vec2 pointA = Mul(xfA, localPoint);
            Vec2 pointB = Mul(xfB, proxyB->vertex(indexB));
            return Dot(pointB - pointA, normal);
        }
        Vec2 normal = Mul(xfB.q, axis);
        Vec2 pointB = Mul(xfB, localPoint);
        Vec2 pointA = Mul(xfA, proxyA->vertex(indexA));
        return Dot(pointA - pointB, normal);
    }
};

enum { kToiUnknown = 0, kToiFailed = 1, kToiOverlapped = 2, kToiTouching = 3, kToiSeparated = 4 };

// b2TimeOfImpact (b2TimeOfImpact.cpp:264-497).  Returns the state, *tOut = output.t.
__device__ __noinline__ int TimeOfImpact(const DistanceProxy* proxyA, const DistanceProxy* proxyB, const Sweep* sweepAIn,
                                         const Sweep* sweepBIn, float tMax, float* tOut) {
    int state = kToiUnknown;
    float tResult = tMax;
    Sweep sweepA = *sweepAIn, sweepB = *sweepBIn;
    SweepNormalize(sweepA);
    SweepNormalize(sweepB);
    const float totalRadius = proxyA->radius + proxyB->radius;
    const float target = Max(kLinearSlop, totalRadius - 3.0f * kLinearSlop);
    const float tolerance = 0.25f * kLinearSlop;
    float t1 = 0.0f;
    const int k_maxIterations = 20;
    int iter = 0;
    SimplexCache cache;
    cache.count = 0;
    cache.metric = 0.0f;
    for (;;) {
        Xf xfA = SweepTransform(sweepA, t1), xfB = SweepTransform(sweepB, t1);
        float distance = GjkDistance(&cache, proxyA, xfA, proxyB, xfB);
        if (distance <= 0.0f) {  // overlapped: give up on continuous collision
            state = kToiOverlapped;
            tResult = 0.0f;
            break;
        }
        if (distance < target + tolerance) {
            state = kToiTouching;
            tResult = t1;
            break;
        }
        SeparationFunction fcn;
        fcn.initialize(&cache, proxyA, sweepA, proxyB, sweepB, t1);
        bool done = false;
        float t2 = tMax;
        int pushBackIter = 0;
        for (;;) {
            int indexA, indexB;
            float s2 = fcn.findMinSeparation(&indexA, &indexB, t2);
            if (s2 > target + tolerance) {
                state = kToiSeparated;
                tResult = tMax;
                done = true;
                break;
            }
            if (s2 > target - tolerance) {
                t1 = t2;
                break;
            }
            float s1 = fcn.evaluate(indexA, indexB, t1);
            if (s1 < target - tolerance) {
                state = kToiFailed;
                tResult = t1;
                done = true;
                break;
            }
            if (s1 <= target + tolerance) {
                state = kToiTouching;
                tResult = t1;
                done = true;
                break;
            }
            int rootIterCount = 0;
            float a1 = t1, a2 = t2;
            for (;;) {
                float t;
                if (rootIterCount & 1) t = a1 + Div((target - s1) * (a2 - a1), s2 - s1);  // secant
                else t = 0.5f * (a1 + a2);                                                 // bisection
                ++rootIterCount;
                float s = fcn.evaluate(indexA, indexB, t);
                if (Abs(s - target) < tolerance) {
                    t2 = t;
                    break;
                }
                if (s > target) {
                    a1 = t;
                    s1 = s;
                } else {
                    a2 = t;
                    s2 = s;
                }
                if (rootIterCount == 50) break;
            }
            ++pushBackIter;
            if (pushBackIter == kMaxPolygonVertices) break;
        }
        ++iter;
        if (done) break;
        if (iter == k_maxIterations) {
            state = kToiFailed;
            tResult = t1;
            break;
        }
    }
    *tOut = tResult;
    return state;
}

// ---- body helpers ----------------------------------------------------------------------------------
struct ToiCols {
    Col<int4> cA;
    Col<int2> cS;
    Col<float> cToi;
    Col<int> bodyFlags, proxyInfo;
    Col<float4> bodyXf, bodySweep, bodySweep0, bodyVel, bodyMass;
};
__device__ __forceinline__ ToiCols MakeToiCols(const Arena& A, int tile, int lane) {
    ToiCols c{col(A.contactA, tile, A.NC, lane), col(A.contactS, tile, A.NC, lane), col(A.contactToi, tile, A.NC, lane),
              col(A.bodyFlags, tile, A.NB, lane), col(A.proxyInfo, tile, A.NP, lane),
              col(A.bodyXf, tile, A.NB, lane), col(A.bodySweep, tile, A.NB, lane), col(A.bodySweep0, tile, A.NB, lane),
              col(A.bodyVel, tile, A.NB, lane), col(A.bodyMass, tile, A.NB, lane)};
    return c;
}
__device__ __forceinline__ Sweep LoadSweep(const ToiCols& c, int b) {
    const float4 s0 = c.bodySweep0[b], s1 = c.bodySweep[b], ms = c.bodyMass[b];
    Sweep s;
    s.localCenter = V2(ms.z, ms.w);
    s.c0 = V2(s0.x, s0.y);
    s.a0 = s0.z;
    s.alpha0 = s0.w;
    s.c = V2(s1.x, s1.y);
    s.a = s1.z;
    return s;
}
// b2Body::SynchronizeTransform (b2Body.h:835-839)
__device__ __forceinline__ void SynchronizeTransform(const ToiCols& c, int b) {
    const float4 sw = c.bodySweep[b], ms = c.bodyMass[b];
    Rot q = MakeRot(sw.z);
    Vec2 p = V2(sw.x, sw.y) - Mul(q, V2(ms.z, ms.w));
    c.bodyXf[b] = make_float4(p.x, p.y, q.s, q.c);
}
// b2Body::Advance (b2Body.h:841-849)
__device__ __forceinline__ void BodyAdvance(const ToiCols& c, int b, float alpha) {
    Sweep s = LoadSweep(c, b);
    SweepAdvance(s, alpha);
    s.c = s.c0;
    s.a = s.a0;
    c.bodySweep0[b] = make_float4(s.c0.x, s.c0.y, s.a0, s.alpha0);
    float4 sw = c.bodySweep[b];
    sw.x = s.c.x; sw.y = s.c.y; sw.z = s.a;
    c.bodySweep[b] = sw;
    Rot q = MakeRot(s.a);
    Vec2 p = s.c - Mul(q, s.localCenter);
    c.bodyXf[b] = make_float4(p.x, p.y, q.s, q.c);
}

// Is contact (flags, bodies) a TOI candidate, per b2World::SolveTOI's filter (b2World.cpp:634-659)?
__device__ __forceinline__ bool ToiCandidate(int bfA, int bfB) {
    const int typeA = bfA & kBodyTypeMask, typeB = bfB & kBodyTypeMask;
    const bool activeA = (bfA & kBodyAwake) && typeA != B2G_STATIC_BODY;
    const bool activeB = (bfB & kBodyAwake) && typeB != B2G_STATIC_BODY;
    if (!activeA && !activeB) return false;
    const bool collideA = (bfA & kBodyBullet) || typeA != B2G_DYNAMIC_BODY;
    const bool collideB = (bfB & kBodyBullet) || typeB != B2G_DYNAMIC_BODY;
    return collideA || collideB;
}

// The part of that filter that cannot change during a step (body types and bullet flags): a contact that fails it
// can never become a candidate, can never join a TOI mini-island (b2World.cpp:806-811 admits only static, kinematic
// or bullet neighbours) and never carries e_toiFlag.  Contacts that pass it form the per-world candidate list.
__device__ __forceinline__ bool ToiStaticCandidate(int bfA, int bfB) {
    return (bfA & kBodyBullet) || (bfA & kBodyTypeMask) != B2G_DYNAMIC_BODY || (bfB & kBodyBullet) ||
           (bfB & kBodyTypeMask) != B2G_DYNAMIC_BODY;
}

// Time of impact of contact k given both sweeps (already on the same alpha0): b2World.cpp:680-709
__device__ __forceinline__ float ContactToi(const Arena& A, int tile, int lane, const ToiCols& c, int k, const int4 a,
                                            const Sweep& sA, const Sweep& sB, float alpha0) {
    const int2 so = c.cS[k];
    const int shapeTypes = (a.w >> 16) & 0xf;
    const int childA = (c.proxyInfo[a.x & 0xffff] >> 16) & 0xffff, childB = (c.proxyInfo[(a.x >> 16) & 0xffff] >> 16) & 0xffff;
    const DistanceProxy pA = MakeDistanceProxy(A, tile, lane, so.x, shapeTypes & 3, childA);
    const DistanceProxy pB = MakeDistanceProxy(A, tile, lane, so.y, (shapeTypes >> 2) & 3, childB);
    float beta;
    const int state = TimeOfImpact(&pA, &pB, &sA, &sB, 1.0f, &beta);
    return state == kToiTouching ? Min(alpha0 + (1.0f - alpha0) * beta, 1.0f) : 1.0f;
}

// b2Contact::Update with its side effects, as called inside SolveTOI (b2Contact.cpp:170-256)
__device__ __forceinline__ void ToiUpdateContact(const Arena& A, int tile, int lane, const ToiCols& c, const EventSink& ev, int k) {
    const int4 a = c.cA[k];
    const int res = UpdateOne(A, tile, lane, k, a);
    const bool wasTouching = (res & kResWasTouching) != 0, touching = (res & kResTouching) != 0;
    const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
    if (touching != wasTouching) {
        WakeBody(c.bodyFlags, c.bodySweep, bodyA);
        WakeBody(c.bodyFlags, c.bodySweep, bodyB);
        A.topoDirty[(size_t)tile * kTile + lane] = 1;  // the islands depend on the touching flags
    }
    if (A.eventMask) {
        const int pointCount = (res >> kResCountShift) & 3;
        const int pa = c.proxyInfo[a.x & 0xffff], pb = c.proxyInfo[(a.x >> 16) & 0xffff];
        const int fA = pa & 0xffff, chA = (pa >> 16) & 0xffff, fB = pb & 0xffff, chB = (pb >> 16) & 0xffff;
        if (!wasTouching && touching) ev.push(kEventBegin, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
        if (wasTouching && !touching) ev.push(kEventEnd, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
        if (touching) ev.push(kEventPreSolve, fA, chA, fB, chB, pointCount, make_float4(0, 0, 0, 0));
    }
}

// The reference's SolveTOI loop for one world, entered after phase 1 has cached every first-pass time of impact.
#ifdef B2G_TOI_PROFILE
// Diagnostic build only (tools/toi_profile.py): SM-clock cycles per section of the event loop, summed over the
// first active lane of every warp.
__device__ unsigned long long g_toiProf[16];
#define TOI_T0() long long t0__ = clock64()
#define TOI_T(slot)                                                                                  \
    do {                                                                                             \
        long long t1__ = clock64();                                                                  \
        if ((__activemask() & ((1u << lane) - 1)) == 0) {                                            \
            atomicAdd(&g_toiProf[slot], (unsigned long long)(t1__ - t0__));                          \
            atomicAdd(&g_toiProf[8 + slot], 1ull);                                                   \
        }                                                                                            \
        t0__ = clock64();                                                                            \
    } while (0)
#define TOI_TA() long long ta__ = clock64()
#define TOI_TB(slot)                                                                                 \
    do {                                                                                             \
        long long tb__ = clock64();                                                                  \
        if ((__activemask() & ((1u << lane) - 1)) == 0) {                                            \
            atomicAdd(&g_toiProf[slot], (unsigned long long)(tb__ - ta__));                          \
            atomicAdd(&g_toiProf[8 + slot], 1ull);                                                   \
        }                                                                                            \
    } while (0)
#else
#define TOI_T0()
#define TOI_T(slot)
#define TOI_TA()
#define TOI_TB(slot)
#endif

constexpr int kToiScan = 8;      // list entries fetched ahead per walk step
constexpr int kToiMaxMoved = 32;  // per world and sub-step, else the ordered fallback
constexpr int kToiMaxCand = 96;

// Per-lane scratch lists of one tile in shared memory (column `lane` belongs to this lane's world).
struct ToiLists {
    uint16_t moved[kToiMaxMoved][kTile];
    uint32_t cand[kToiMaxCand][kTile];
    unsigned char exists[kToiMaxCand][kTile];
};

// b2BroadPhase::UpdatePairs + b2ContactManager::AddPair for one world by one lane, after a TOI sub-step
// (b2World.cpp:883-885).  Same result as UpdatePairsOrdered; organised like k_sync_pairs so that an existing
// contact is found by ONE walk over the contact array for all candidate pairs instead of one walk per pair:
// moved proxies -> overlap candidates (min, max) -> sort + dedupe -> mark the pairs that already have a contact
// -> AddPair for the rest in (proxyIdA, proxyIdB) order.
__device__ __noinline__ void UpdatePairsLane(const Arena* Ap, int tile, int lane, ToiLists* L) {
    const Arena& A = *Ap;
    const size_t ws = (size_t)tile * kTile + lane;
    const int np = A.proxyCount[ws];
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    int nMoved = 0;
    for (int p0 = 0; p0 < np; p0 += kToiScan) {
        int mv[kToiScan];
#pragma unroll
        for (int j = 0; j < kToiScan; ++j) mv[j] = p0 + j < np ? proxyMoved[p0 + j] : 0;
#pragma unroll
        for (int j = 0; j < kToiScan; ++j)
            if (mv[j]) {
                if (nMoved < kToiMaxMoved) L->moved[nMoved][lane] = (uint16_t)(p0 + j);
                ++nMoved;
            }
    }
    if (nMoved == 0) return;
    if (nMoved > kToiMaxMoved) { UpdatePairsOrdered(Ap, tile, lane); return; }
    int nCand = 0;
    for (int m = 0; m < nMoved; ++m) {
        const int pm = L->moved[m][lane];
        const float4 fatM = proxyFat[pm];
        for (int q0 = 0; q0 < np; q0 += kToiScan) {
            float4 fq[kToiScan];
#pragma unroll
            for (int j = 0; j < kToiScan; ++j)
                if (q0 + j < np) fq[j] = proxyFat[q0 + j];
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) {
                const int q = q0 + j;
                if (q >= np) break;
                if (q == pm || !FatOverlap(fatM, fq[j])) continue;
                if (nCand < kToiMaxCand) L->cand[nCand][lane] = q < pm ? ((uint32_t)q << 16 | (uint32_t)pm) : ((uint32_t)pm << 16 | (uint32_t)q);
                ++nCand;
            }
        }
    }
    if (nCand > kToiMaxCand) { UpdatePairsOrdered(Ap, tile, lane); return; }  // the move flags are still set
    for (int m = 0; m < nMoved; ++m) proxyMoved[L->moved[m][lane]] = 0;  // the move buffer is consumed
    if (nCand == 0) return;
    for (int i = 1; i < nCand; ++i) {  // std::sort by (proxyIdA, proxyIdB); the lists are short
        const uint32_t v = L->cand[i][lane];
        int j = i - 1;
        while (j >= 0 && L->cand[j][lane] > v) {
            L->cand[j + 1][lane] = L->cand[j][lane];
            --j;
        }
        L->cand[j + 1][lane] = v;
    }
    int u = 0;
    for (int i = 0; i < nCand; ++i) {  // the duplicate skip of UpdatePairs
        const uint32_t v = L->cand[i][lane];
        if (u > 0 && L->cand[u - 1][lane] == v) continue;
        L->cand[u][lane] = v;
        L->exists[u][lane] = 0;
        ++u;
    }
    nCand = u;
    PairCtx c = MakePairCtx(A, tile, lane);
    for (int k0 = 0; k0 < c.nc; k0 += kToiScan) {  // which candidates already have a contact
        uint32_t xs[kToiScan];
#pragma unroll
        for (int j = 0; j < kToiScan; ++j) xs[j] = k0 + j < c.nc ? (uint32_t)c.cA[k0 + j].x : 0xffffffffu;
        for (int i = 0; i < nCand; ++i) {
            const uint32_t v = L->cand[i][lane], vs = (v >> 16) | (v << 16);
            bool hit = false;
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) hit |= xs[j] == v || xs[j] == vs;
            if (hit) L->exists[i][lane] = 1;
        }
    }
    for (int i = 0; i < nCand; ++i) {
        if (L->exists[i][lane]) continue;
        const uint32_t v = L->cand[i][lane];
        AddPair(c, (int)(v >> 16), (int)(v & 0xffff), true);
    }
    if (c.nc != c.ncAtStart) A.contactCount[ws] = c.nc;
}

__device__ __noinline__ void SolveToiWorld(const Arena* Ap, int tile, int lane, float dt, int velIters, ToiLists* lists) {
    const Arena& A = *Ap;
    const size_t ws = (size_t)tile * kTile + lane;
    const ToiCols c = MakeToiCols(A, tile, lane);
    Col<int> toiBodies = col(A.toiBodies, tile, kToiBodyCap, lane);
    Col<int> toiContacts = col(A.toiContacts, tile, kMaxToiContacts, lane);
    Col<float4> vc = col(A.toiVc, tile, kMaxToiContacts * kVcStride, lane);
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    EventSink ev{col(A.eventA, tile, A.NE, lane), col(A.eventB, tile, A.NE, lane), &A.eventCount[ws], &A.status[ws],
                 A.NE, A.eventMask};
    // contact indices that pass ToiStaticCandidate, ascending (k_toi); the world's contact list is that array
    // walked backwards, so every list walk below visits the candidates from the end of this list to its start
    Col<int> candList = col(A.toiCand, tile, A.NC, lane);
    int nCandList = A.toiCandCount[ws];
    Col<int> bodyProxyStart = col(A.bodyProxyStart, tile, A.NB + 1, lane);
    Col<int> bodyProxyList = col(A.bodyProxyList, tile, A.NP, lane);

    for (;;) {
        TOI_T0();
        // ---- find the first TOI (b2World.cpp:598-722): world contact list order = descending array index ----
        const int n = A.contactCount[ws];
        int minContact = -1;
        float minAlpha = 1.0f;
        // The list walks of this loop are chains of dependent global loads when written naively (load, branch,
        // next); each walk fetches kToiScan contacts ahead into registers so that one memory latency covers
        // kToiScan list entries.  No entry other than the one being processed is written inside a walk.
        for (int i0 = nCandList - 1; i0 >= 0; i0 -= kToiScan) {
            int kk[kToiScan];
            int4 aa[kToiScan];
            float tt[kToiScan];
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) kk[j] = i0 - j >= 0 ? candList[i0 - j] : 0;
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) {
                if (i0 - j >= 0) {
                    aa[j] = c.cA[kk[j]];
                    tt[j] = c.cToi[kk[j]];
                }
            }
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) {
                if (i0 - j < 0) break;
                const int k = kk[j];
                int4 a = aa[j];
                if ((a.z & kContactEnabled) == 0) continue;
                if (((a.z & kContactToiCountMask) >> kContactToiCountShift) > kMaxSubSteps) continue;
                float alpha = 1.0f;
                if (a.z & kContactToi) {
                    alpha = tt[j];
                } else {
                    const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
                    if (!ToiCandidate(c.bodyFlags[bodyA], c.bodyFlags[bodyB])) continue;
                    // put the sweeps onto the same time interval
                    Sweep sA = LoadSweep(c, bodyA), sB = LoadSweep(c, bodyB);
                    float alpha0 = sA.alpha0;
                    if (sA.alpha0 < sB.alpha0) {
                        alpha0 = sB.alpha0;
                        SweepAdvance(sA, alpha0);
                        c.bodySweep0[bodyA] = make_float4(sA.c0.x, sA.c0.y, sA.a0, sA.alpha0);
                    } else if (sB.alpha0 < sA.alpha0) {
                        alpha0 = sA.alpha0;
                        SweepAdvance(sB, alpha0);
                        c.bodySweep0[bodyB] = make_float4(sB.c0.x, sB.c0.y, sB.a0, sB.alpha0);
                    }
                    TOI_TA();
                    alpha = ContactToi(A, tile, lane, c, k, a, sA, sB, alpha0);
                    TOI_TB(6);
                    c.cToi[k] = alpha;
                    a.z |= kContactToi;
                    c.cA[k] = a;
                }
                if (alpha < minAlpha) {
                    minContact = k;
                    minAlpha = alpha;
                }
            }
        }
        TOI_T(0);
        if (minContact < 0 || 1.0f - 10.0f * kEpsilon < minAlpha) break;  // no more TOI events

        // topoDirty (k_islands' cache key) is raised only by what the islands depend on: a touching flag that
        // flips (ToiUpdateContact), a contact that is disabled or created (below, AddPair), a body that wakes.
        // ---- advance the bodies to the TOI (b2World.cpp:730-760) ----
        int4 mc = c.cA[minContact];
        const int bA = mc.y & 0xffff, bB = (mc.y >> 16) & 0xffff;
        const float4 backupSwA = c.bodySweep[bA], backupSw0A = c.bodySweep0[bA];
        const float4 backupSwB = c.bodySweep[bB], backupSw0B = c.bodySweep0[bB];
        BodyAdvance(c, bA, minAlpha);
        BodyAdvance(c, bB, minAlpha);
        ToiUpdateContact(A, tile, lane, c, ev, minContact);
        mc = c.cA[minContact];
        mc.z &= ~kContactToi;
        mc.z += 1 << kContactToiCountShift;
        if ((mc.z & kContactEnabled) == 0 || (mc.z & kContactTouching) == 0) {  // not solid: restore the sweeps
            mc.z &= ~kContactEnabled;
            c.cA[minContact] = mc;
            A.topoDirty[ws] = 1;
            float4 sw = backupSwA;
            sw.w = c.bodySweep[bA].w;  // m_sleepTime is not part of b2Sweep
            c.bodySweep[bA] = sw;
            c.bodySweep0[bA] = backupSw0A;
            sw = backupSwB;
            sw.w = c.bodySweep[bB].w;
            c.bodySweep[bB] = sw;
            c.bodySweep0[bB] = backupSw0B;
            SynchronizeTransform(c, bA);
            SynchronizeTransform(c, bB);
            continue;
        }
        if (WakeBody(c.bodyFlags, c.bodySweep, bA) | WakeBody(c.bodyFlags, c.bodySweep, bB)) A.topoDirty[ws] = 1;
        TOI_T(1);

        // ---- build the island (b2World.cpp:765-857) ----
        int nB = 0, nC = 0;
        toiBodies[nB++] = bA;
        toiBodies[nB++] = bB;
        toiContacts[nC++] = minContact;
        c.bodyFlags[bA] |= kBodyToiIsland;
        c.bodyFlags[bB] |= kBodyToiIsland;
        mc.z |= kContactToiIsland;
        c.cA[minContact] = mc;
        for (int i = 0; i < 2; ++i) {
            const int body = i == 0 ? bA : bB;
            if ((c.bodyFlags[body] & kBodyTypeMask) != B2G_DYNAMIC_BODY) continue;
            // the body's contact edge list is head-inserted: newest contact (highest array index) first
            bool full = false;
            for (int i0 = nCandList - 1; i0 >= 0 && !full; i0 -= kToiScan) {
                int kk[kToiScan];
                int4 aa[kToiScan];
#pragma unroll
                for (int j = 0; j < kToiScan; ++j) kk[j] = i0 - j >= 0 ? candList[i0 - j] : 0;
#pragma unroll
                for (int j = 0; j < kToiScan; ++j)
                    if (i0 - j >= 0) aa[j] = c.cA[kk[j]];
#pragma unroll
                for (int j = 0; j < kToiScan; ++j) {
                    if (i0 - j < 0) break;
                    const int k = kk[j];
                    int4 a = aa[j];
                    const int kA = a.y & 0xffff, kB = (a.y >> 16) & 0xffff;
                    if (kA != body && kB != body) continue;
                    if (nB == kToiBodyCap || nC == kMaxToiContacts) { full = true; break; }
                    if (a.z & kContactToiIsland) continue;
                    const int other = kA == body ? kB : kA;
                    const int of = c.bodyFlags[other];
                    // only static, kinematic or bullet bodies
                    if ((of & kBodyTypeMask) == B2G_DYNAMIC_BODY && (c.bodyFlags[body] & kBodyBullet) == 0 && (of & kBodyBullet) == 0)
                        continue;
                    // tentatively advance the body to the TOI
                    const float4 backupSw = c.bodySweep[other], backupSw0 = c.bodySweep0[other];
                    if ((of & kBodyToiIsland) == 0) BodyAdvance(c, other, minAlpha);
                    ToiUpdateContact(A, tile, lane, c, ev, k);
                    a = c.cA[k];
                    if ((a.z & kContactEnabled) == 0 || (a.z & kContactTouching) == 0) {
                        float4 sw = backupSw;
                        sw.w = c.bodySweep[other].w;  // m_sleepTime is not part of b2Sweep
                        c.bodySweep[other] = sw;
                        c.bodySweep0[other] = backupSw0;
                        SynchronizeTransform(c, other);
                        continue;
                    }
                    a.z |= kContactToiIsland;
                    c.cA[k] = a;
                    toiContacts[nC++] = k;
                    const int of2 = c.bodyFlags[other];
                    if (of2 & kBodyToiIsland) continue;
                    c.bodyFlags[other] = of2 | kBodyToiIsland;
                    if ((of2 & kBodyTypeMask) != B2G_STATIC_BODY && WakeBody(c.bodyFlags, c.bodySweep, other)) A.topoDirty[ws] = 1;
                    toiBodies[nB++] = other;
                }
            }
        }

        TOI_T(2);
        // ---- b2Island::SolveTOI (b2Island.cpp:385-515): positions / velocities live in the body arrays ----
        const float h = (1.0f - minAlpha) * dt;
        for (int i = 0; i < nC; ++i) InitVelocityConstraint(A, tile, lane, vc, i, toiContacts[i], false, 1.0f);  // solver ctor
        for (int it = 0; it < 20; ++it) {
            float minSeparation = 0.0f;
            for (int i = 0; i < nC; ++i)
                SolvePositionConstraint(A, tile, lane, vc, i, toiContacts[i], kToiBaumgarte, true, bA, bB, &minSeparation);
            if (minSeparation >= -1.5f * kLinearSlop) break;
        }
        {  // leap of faith to the new safe state
            float4 s = c.bodySweep[bA], s0 = c.bodySweep0[bA];
            c.bodySweep0[bA] = make_float4(s.x, s.y, s.z, s0.w);
            s = c.bodySweep[bB]; s0 = c.bodySweep0[bB];
            c.bodySweep0[bB] = make_float4(s.x, s.y, s.z, s0.w);
        }
        for (int i = 0; i < nC; ++i) InitVelocityConstraint(A, tile, lane, vc, i, toiContacts[i], false, 1.0f);
        for (int it = 0; it < velIters; ++it)
            for (int i = 0; i < nC; ++i) SolveVelocityConstraint(vc, c.bodyVel, i);
        for (int i = 0; i < nB; ++i) {  // integrate positions, sync bodies
            const int b = toiBodies[i];
            float4 sw = c.bodySweep[b];
            const float4 vel = c.bodyVel[b];
            Vec2 cc = V2(sw.x, sw.y);
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
            cc = cc + h * v;
            a += h * wv;
            c.bodySweep[b] = make_float4(cc.x, cc.y, a, sw.w);
            c.bodyVel[b] = make_float4(v.x, v.y, wv, 0.0f);
            SynchronizeTransform(c, b);
        }
        if (A.eventMask >> kEventPostSolve & 1) {  // b2Island::Report
            for (int i = 0; i < nC; ++i) {
                const size_t o = (size_t)i * kVcStride;
                const float4 imp = vc[o + 8];
                const int vcPointCount = __float_as_int(vc[o + 3].w) & 0xff;
                const int4 a = c.cA[toiContacts[i]];
                const int pa = c.proxyInfo[a.x & 0xffff], pb = c.proxyInfo[(a.x >> 16) & 0xffff];
                ev.push(kEventPostSolve, pa & 0xffff, (pa >> 16) & 0xffff, pb & 0xffff, (pb >> 16) & 0xffff, vcPointCount,
                        make_float4(imp.x, vcPointCount > 1 ? imp.z : 0.0f, imp.y, vcPointCount > 1 ? imp.w : 0.0f));
            }
        }

        TOI_T(3);
        // ---- reset island flags, synchronise broad-phase proxies, invalidate TOIs (b2World.cpp:868-886) ----
        for (int i = 0; i < nB; ++i) {
            const int b = toiBodies[i];
            int f = c.bodyFlags[b] & ~kBodyToiIsland;
            if ((f & kBodyTypeMask) == B2G_DYNAMIC_BODY) f |= kBodyToiMoved;
            c.bodyFlags[b] = f;
        }
        for (int i = 0; i < nB; ++i) {  // b2Body::SynchronizeFixtures of the displaced bodies
            const int b = toiBodies[i];
            if ((c.bodyFlags[b] & kBodyToiMoved) == 0) continue;
            const int pEnd = bodyProxyStart[b + 1];
            for (int j = bodyProxyStart[b]; j < pEnd; ++j) {
                const int p = bodyProxyList[j];
                const int pi = c.proxyInfo[p];
                const int fi = fixtureInfo[pi & 0xffff];
                SynchronizeProxy(A, tile, lane, p, pi & 0xffff, (pi >> 16) & 0xffff, fi, b);
            }
        }
        TOI_TA();
        for (int i0 = 0; i0 < nCandList; i0 += kToiScan) {  // only candidates carry e_toiFlag / the TOI island flag
            int kk[kToiScan];
            int4 aa[kToiScan];
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) kk[j] = i0 + j < nCandList ? candList[i0 + j] : 0;
#pragma unroll
            for (int j = 0; j < kToiScan; ++j)
                if (i0 + j < nCandList) aa[j] = c.cA[kk[j]];
#pragma unroll
            for (int j = 0; j < kToiScan; ++j) {
                if (i0 + j >= nCandList) break;
                const int k = kk[j];
                int4 a = aa[j];
                if ((a.z & (kContactToi | kContactToiIsland)) == 0) continue;
                if ((c.bodyFlags[a.y & 0xffff] | c.bodyFlags[(a.y >> 16) & 0xffff]) & kBodyToiMoved) {
                    a.z &= ~(kContactToi | kContactToiIsland);
                    c.cA[k] = a;
                }
            }
        }
        TOI_TB(7);
        for (int i = 0; i < nB; ++i) c.bodyFlags[toiBodies[i]] &= ~kBodyToiMoved;
        TOI_T(4);
        UpdatePairsLane(Ap, tile, lane, lists);  // m_contactManager.FindNewContacts()
        for (int k = n, nNew = A.contactCount[ws]; k < nNew; ++k) {  // new contacts sit at the head of the list
            const int4 a = c.cA[k];
            if ((a.z & kContactSensor) == 0 && ToiStaticCandidate(c.bodyFlags[a.y & 0xffff], c.bodyFlags[(a.y >> 16) & 0xffff]))
                candList[nCandList++] = k;
        }
        TOI_T(5);
    }
}

// =================================================================================================
// k_toi — b2World::SolveTOI for every world with continuous physics
// =================================================================================================
__global__ void __launch_bounds__(kToiWarps * 32) k_toi(Arena A, float dt, int velIters) {
    // [NC][kTile]: contact passes ToiStaticCandidate (k_collide's result-byte scratch, free again at this point of the chain)
    unsigned char* sIsCand = A.collideRes + (size_t)(A.tile0 + blockIdx.x) * A.NC * kTile;
    __shared__ unsigned int sMinAlpha[kTile];
    __shared__ int sItems;
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const bool live = tile * kTile + lane < A.numWorlds && (A.worldFlags[ws] & B2G_WORLD_CONTINUOUS) != 0;
    if (!__syncthreads_or(live)) return;
    if (warp == 0) sMinAlpha[lane] = __float_as_uint(1.0f);
    if (threadIdx.x == 0) sItems = 0;
    __syncthreads();
    const ToiCols c = MakeToiCols(A, tile, lane);
    const int n = live ? A.contactCount[ws] : 0;
    const int nb = live ? A.bodyCount[ws] : 0;

    // ---- b2World.cpp:582-596: clear island flags, alpha0 = 0, invalidate every TOI.  alpha0 is only read by
    // phase 2 (after the barrier); phase 1 knows it is 0. ----
    const int nbMax = __reduce_max_sync(0xffffffffu, nb);
    for (int b = warp; b < nbMax; b += kToiWarps) {
        if (b >= nb) continue;
        float4 s0 = c.bodySweep0[b];
        if (s0.w != 0.0f) {
            s0.w = 0.0f;
            c.bodySweep0[b] = s0;
        }
        const int f = c.bodyFlags[b];
        if (f & (kBodyToiIsland | kBodyToiMoved)) c.bodyFlags[b] = f & ~(kBodyToiIsland | kBodyToiMoved);
    }

    __syncthreads();

    // ---- phase 1: first pass over the contact list, every candidate's time of impact in parallel ----
    // With lane = world, contact k is a candidate in a few of a tile's 32 worlds only (20 of a pyramid's 590 contacts touch
    // the ground, at different indices in every jittered world), so a warp running b2TimeOfImpact where it stands would run
    // it for one or two lanes at a time.  1a: the candidates of the tile are collected as (contact, lane) items; 1b: the
    // warps evaluate 32 items each, every lane busy.  An item only touches its own world's columns.  Measured against
    // evaluating in place: 16,384 mixed-shape worlds 1.34 -> 0.83 ms (their contact counts differ, the items balance the
    // warps); 16,384 pyramids 0.44 -> 0.51 ms (their ground contacts are the first contacts every world creates and line
    // up anyway; the items' scattered loads cost a little).  A run of 32 candidates of one contact index stays one
    // coalesced run of items.
    uint32_t* items = reinterpret_cast<uint32_t*>(A.velRec ? (void*)(A.velRec + (size_t)tile * A.NC * 7 /* kVelRec */ * kTile)
                                                           : (void*)(A.vc + (size_t)tile * A.NC * kVcStride * kTile));  // dead scratch
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    for (int k = warp; k < nMax; k += kToiWarps) {
        if (k >= n) continue;
        int4 a = c.cA[k];
        const int z0 = a.z;
        a.z &= ~(kContactToi | kContactToiIsland | kContactToiCountMask);
        const int bodyA = a.y & 0xffff, bodyB = (a.y >> 16) & 0xffff;
        const int bfA = c.bodyFlags[bodyA], bfB = c.bodyFlags[bodyB];
        const bool sensor = (a.z & kContactSensor) != 0;  // b2World.cpp:632-636: sensors never take part
        sIsCand[k * kTile + lane] = (!sensor && ToiStaticCandidate(bfA, bfB)) ? 1 : 0;
        const bool cand = (a.z & kContactEnabled) && !sensor && ToiCandidate(bfA, bfB);
        const unsigned who = __ballot_sync(__activemask(), cand);
        if (cand) {
            int base = 0;
            const int leader = __ffs(who) - 1;
            if (lane == leader) base = atomicAdd(&sItems, __popc(who));
            base = __shfl_sync(who, base, leader);
            items[base + __popc(who & ((1u << lane) - 1u))] = (uint32_t)k | ((uint32_t)lane << 16);
            a.z |= kContactToi;
        }
        if (a.z != z0) c.cA[k] = a;
    }
    __syncthreads();
    const int total = sItems;
    for (int i = threadIdx.x; i < total; i += kToiWarps * 32) {
        const uint32_t item = items[i];
        const int k = (int)(item & 0xffff), ln = (int)(item >> 16);
        const ToiCols ci = MakeToiCols(A, tile, ln);
        const int4 a = ci.cA[k];
        Sweep sA = LoadSweep(ci, a.y & 0xffff), sB = LoadSweep(ci, (a.y >> 16) & 0xffff);
        sA.alpha0 = 0.0f;
        sB.alpha0 = 0.0f;
        const float alpha = ContactToi(A, tile, ln, ci, k, a, sA, sB, 0.0f);
        ci.cToi[k] = alpha;
        if (alpha < 1.0f) atomicMin(&sMinAlpha[ln], __float_as_uint(alpha));  // alpha >= 0: uint order == float order
    }
    __syncthreads();
    if (warp == 0) {
        const float minAlpha = live ? __uint_as_float(sMinAlpha[lane]) : 1.0f;
        A.toiMinAlpha[ws] = minAlpha;
        if (!(1.0f - 10.0f * kEpsilon < minAlpha)) {  // k_toi_events will run this world: give it the candidate list
            Col<int> candList = col(A.toiCand, tile, A.NC, lane);
            int m = 0;
            for (int k = 0; k < n; ++k)
                if (sIsCand[k * kTile + lane]) candList[m++] = k;
            A.toiCandCount[ws] = m;
        }
    }
}

// k_toi_events — phase 2: worlds with a TOI event run the reference's loop, one world per lane.  A CTA is one
// warp (one tile) so that many tiles are resident per SM: the loop is a chain of dependent global loads and the
// only way to hide them is other tiles' warps.
__global__ void __launch_bounds__(32) k_toi_events(Arena A, float dt, int velIters) {
    __shared__ ToiLists lists;
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x;
    const size_t ws = (size_t)tile * kTile + lane;
    if (tile * kTile + lane >= A.numWorlds) return;
    if (1.0f - 10.0f * kEpsilon < A.toiMinAlpha[ws]) return;
    SolveToiWorld(&A, tile, lane, dt, velIters, &lists);
}

}  // namespace b2g
