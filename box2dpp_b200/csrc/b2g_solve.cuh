// b2g_solve.cuh — b2Island::Solve + b2ContactSolver as a chain of sm_100a kernels (the "tile" path).
//
// Why several kernels: inside one world the sequential-impulse solver is one dependency chain —
// consecutive island contacts share a body (the island DFS emits a body's contacts back to back), so
// Gauss-Seidel order leaves no contact-level parallelism.  The chain is split from the
// work around it:
//
//   k_solve_prepare    parallel   integrate velocities (b2Island.cpp:190-225), b2ContactSolver ctor +
//                                 InitializeVelocityConstraints (b2ContactSolver.cpp:33-238) -> record streams
//   k_solve_velocity   SEQUENTIAL WarmStart + velocity iterations (b2ContactSolver.cpp:240-597), one warp per
//                                 tile of 32 worlds, lane = world; body velocities of the tile live in shared
//                                 memory (TMA bulk copy in / out); constraint records stream through a
//                                 TMA + mbarrier ring; accumulated impulses through a cp.async ring
//   k_solve_integrate  parallel   StoreImpulses (:599-612), PostSolve report (b2Island.cpp:517-540), integrate
//                                 positions (b2Island.cpp:276-306)
//   k_solve_position   SEQUENTIAL position iterations with per-island early exit (b2Island.cpp:308-328,
//                                 b2ContactSolver.cpp:614-744); body positions in shared memory
//   k_solve_finalize   parallel   write-back + SynchronizeTransform (b2Island.cpp:331-339), sleep (:345-382)
//
// Islands of one world are independent (they share only static bodies, which never change), so every phase
// runs over the concatenation of all islands of a world; per island the order of operations — and therefore
// every rounding — is exactly the reference's.
#pragma once

#include <type_traits>

#include "b2g_kernels.cuh"

namespace b2g {

constexpr int kVelRec = 7;      // float4 per velocity record
constexpr int kPosRec = 5;      // float4 per position record
#ifndef B2G_REC_STAGES
#define B2G_REC_STAGES 8
#endif
constexpr int kRecStages = B2G_REC_STAGES;   // TMA ring depth of the position kernel (records in flight per tile)
#ifndef B2G_VEL_STAGES
#define B2G_VEL_STAGES 8
#endif
constexpr int kVelStages = B2G_VEL_STAGES;   // TMA ring depth of the velocity kernel
#ifndef B2G_VEL_PREFETCH
#define B2G_VEL_PREFETCH 0
#endif
#ifndef B2G_VEL_PIPELINE_MIN
#define B2G_VEL_PIPELINE_MIN 32
#endif
constexpr int kVelPipelineMin = B2G_VEL_PIPELINE_MIN;              // island contacts per world from which the velocity kernel pipelines
// Records pulled into L2 ahead of the rings' own loads (cp.async.bulk.prefetch.L2).  OFF: it bought nothing once the ring
// released its stages early (velocity kernel 1.796 ms without, 1.786 ms with, profiles/r2h_*), and with it compiled in two
// runs of the same 16,384-world batch disagreed in a few whole tiles per 120 steps when two world groups ran side by side
// (tools/diag_determinism.py; never without it).  Kept as a build knob for that experiment only.
constexpr int kVelPrefetch = B2G_VEL_PREFETCH;
#ifndef B2G_POS_PREFETCH
#define B2G_POS_PREFETCH 0
#endif
constexpr int kPosPrefetch = B2G_POS_PREFETCH;
#ifndef B2G_POS_FORCE_SIMPLE
#define B2G_POS_FORCE_SIMPLE 0
#endif
// A/B knob: give a ring stage back to the producer as soon as its record has been LOADED (1) instead of once it has
// been SOLVED (0, default).  Early release is not safe here: mbarrier.arrive is not held back by shared-memory loads that
// are still in flight, and nothing else is either when the loaded registers are first used a whole solve later.  With
// another kernel hammering the same SM's load / store unit (two world groups side by side) a record's loads now and then
// came back after the producer had refilled the stage: runs of 8 lanes (one 128-byte wavefront of a float4 row) holding
// the record eight places further on, i.e. two runs of one batch differing in a few tiles per thousand steps
// (tools/diag_determinism.py, profiles/r2_determinism.md).  Released after the solve, every loaded register has been
// consumed by the arithmetic whose results the preceding st.shared wrote, so the loads have returned by construction;
// the two stages of look-ahead this gives up cost nothing measurable (velocity kernel 1.733 -> 1.744 ms).
#ifndef B2G_EARLY_RELEASE
#define B2G_EARLY_RELEASE 0
#endif
constexpr int kImpStages = 8;   // cp.async ring depth for accumulated impulses
constexpr int kStageBytes = kVelRec * 16 * kTile;  // 3584
constexpr int kPrepWarps = 8;

// ---- PTX: mbarrier + 1-D TMA bulk copies + cp.async ----------------------------------------------
__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void MbarInit(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(SmemAddr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void TmaLoad1D(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(SmemAddr(dst)),
                 "l"(src), "r"(bytes), "r"(SmemAddr(bar))
                 : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void TmaStore1D(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(SmemAddr(src)), "r"(bytes)
                 : "memory");
}
// global -> L2 bulk prefetch (no destination: the later TMA load of the same bytes finds them in L2)
__device__ __forceinline__ void TmaPrefetchL2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void TmaStoreCommitAndWait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void FenceProxyAsync() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void FenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void CpAsync16(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(SmemAddr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void CpAsyncWait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Roles of the sequential kernels' four warps.  Two of these CTAs share an SM, and what matters is that their two consumer
// warps get a warp scheduler each.  Measured (tools/micro/smsp_map.cu, profiles/r2_smsp_map.txt): the hardware hands the
// four warp slots of a CTA out rotated, differently for the two CTAs of an SM (warps 0..3 -> slots 2 3 0 1 and 7 4 5 6),
// and slot % 4 is the scheduler; so the SAME warp index in both CTAs never shares a scheduler, while warp w of one and
// w - 1 of the other do.  (The three-warp CTAs of round 1 had their consumers on one scheduler: 2.38 ms instead of
// 1.78 ms for the velocity kernel.)  Hence: four warps, fixed roles, consumer = warp 0 everywhere.
enum SeqRole { kRoleConsumer = 0, kRoleIdle = 1, kRoleProducer = 2, kRoleIdle2 = 3 };
__device__ __forceinline__ int SequentialRole(int warp) { return warp; }

__device__ __forceinline__ int PackBodies(int bodyA, int bodyB, int pointCount) { return bodyA | (bodyB << 15) | (pointCount << 30); }

// Shared-memory budget of the sequential kernels for a world capacity of NB bodies.
// The tile's body table in shared memory: NB rows of 3 floats x 32 lanes, plus one scratch row that lanes without a
// contact at the current index work on.
__host__ __device__ inline size_t SolveTableBytes(int NB) { return ((size_t)(NB + 1) * 3 * kTile * 4 + 127) & ~(size_t)127; }
__host__ __device__ inline size_t SolveSmemBytes(int NB) {
    size_t body = SolveTableBytes(NB);
    size_t done = 2 * (size_t)((NB + 31) / 32) * kTile * 4;  // "converged" and "contacts fine this pass" bit per island
    size_t vel = (size_t)kVelStages * (kVelRec + 1) * 16 * kTile + 256;                // ring + barriers
    size_t pos = (size_t)kRecStages * kPosRec * 16 * kTile + done + 256;
    return body + (vel > pos ? vel : pos);
}

// =================================================================================================
// k_solve_prepare: CTA = one tile, kPrepWarps warps; warp w takes items w, w + kPrepWarps, ...
// =================================================================================================
#ifndef B2G_PREP_MIN_BLOCKS
#define B2G_PREP_MIN_BLOCKS 4  // 64 registers: measured best of 1..6 (0.58 -> 0.55 ms)
#endif
__global__ void __launch_bounds__(kPrepWarps * 32, B2G_PREP_MIN_BLOCKS) k_solve_prepare(Arena A, float h) {
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const bool live = tile * kTile + lane < A.numWorlds;
    const int nb = live ? A.bodyCount[ws] : 0;
    const int islands = live ? A.islandCount[ws] : 0;
    const int worldFlags = A.worldFlags[ws];
    const bool warmStarting = (worldFlags & B2G_WORLD_WARM_STARTING) != 0;
    const bool clearForces = (worldFlags & B2G_WORLD_CLEAR_FORCES) != 0;
    const float dtRatio = A.invDt0[ws] * h;
    const float2 g2 = A.gravity[ws];
    const Vec2 gravity = V2(g2.x, g2.y);

    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    Col<float4> bodyForce = col(A.bodyForce, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float4> bodyDamp = col(A.bodyDamp, tile, A.NB, lane);
    float* sv = A.solveVel + (size_t)tile * A.NB * 3 * kTile + lane;  // [body][3][lane]
    if (warp == 0) A.solveFlags[ws] = 0;  // set again below, behind the barrier

    // ---- integrate velocities of island bodies; every body's velocity goes into the solver table ----
    const int nbMax = __reduce_max_sync(0xffffffffu, nb);
    for (int b = warp; b < nbMax; b += kPrepWarps) {
        if (b >= nb) continue;
        const int f = bodyFlags[b];
        float4 vel = bodyVel[b];
        if (f & kBodyIsland) {  // set by k_islands on every non-static island body
            const float4 sw = bodySweep[b];
            bodySweep0[b] = make_float4(sw.x, sw.y, sw.z, 0.0f);
            if ((f & kBodyTypeMask) == B2G_DYNAMIC_BODY) {
                const float4 frc = bodyForce[b];
                const float4 ms = bodyMass[b];
                const float4 dp = bodyDamp[b];
                Vec2 v = V2(vel.x, vel.y);
                float wv = vel.z;
                v = v + h * (dp.z * gravity + ms.x * V2(frc.x, frc.y));
                wv += h * ms.y * frc.z;
                v = Div(1.0f, 1.0f + h * dp.x) * v;
                wv *= Div(1.0f, 1.0f + h * dp.y);
                vel = make_float4(v.x, v.y, wv, 0.0f);
                if (clearForces && (frc.x != 0.0f || frc.y != 0.0f || frc.z != 0.0f))
                    bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        sv[(size_t)(b * 3 + 0) * kTile] = vel.x;
        sv[(size_t)(b * 3 + 1) * kTile] = vel.y;
        sv[(size_t)(b * 3 + 2) * kTile] = vel.z;
    }
    __syncthreads();

    // ---- contact records ----
    Col<int> iContacts = col(A.islandContacts, tile, A.NC, lane);
    Col<int> iContactEnd = col(A.islandContactEnd, tile, A.NB, lane);
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<int4> cK = col(A.contactK, tile, A.NC, lane);
    Col<int2> cS = col(A.contactS, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> cN = col(A.contactN, tile, A.NC, lane);
    float4* velRec = A.velRec + (size_t)tile * A.NC * kVelRec * kTile + lane;
    float4* posRec = A.posRec + (size_t)tile * A.NC * kPosRec * kTile + lane;
    float4* impRec = A.impRec + (size_t)tile * A.NC * kTile + lane;

    const int n = islands > 0 ? iContactEnd[islands - 1] : 0;
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    if (warp == 0) A.solveCount[ws] = n;
    int isl = 0, islEnd = islands > 0 ? iContactEnd[0] : 0;  // island of contact i (i only grows)
    bool circles = false;
    for (int i = warp; i < nMax; i += kPrepWarps) {
        if (i >= n) continue;
        while (i >= islEnd) islEnd = iContactEnd[++isl];
        const int k = iContacts[i];
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
        const Vec2 cAp = V2(swA.x, swA.y), cBp = V2(swB.x, swB.y);
        const Vec2 vA = V2(sv[(size_t)(bodyA * 3 + 0) * kTile], sv[(size_t)(bodyA * 3 + 1) * kTile]);
        const Vec2 vB = V2(sv[(size_t)(bodyB * 3 + 0) * kTile], sv[(size_t)(bodyB * 3 + 1) * kTile]);
        const float wA = sv[(size_t)(bodyA * 3 + 2) * kTile], wB = sv[(size_t)(bodyB * 3 + 2) * kTile];
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
        float4 r2, r3, r4 = make_float4(0, 0, 0, 0), r5 = make_float4(0, 0, 0, 0), r6 = make_float4(0, 0, 0, 0);
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
            r2 = make_float4(rA0.x, rA0.y, rB0.x, rB0.y);
            r3 = make_float4(normalMass, tangentMass, velocityBias, 0.0f);
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
            r4 = make_float4(rA1.x, rA1.y, rB1.x, rB1.y);
            r3.w = velocityBias;
            // block solver preparation
            float rn1A = Cross(rA0, normal), rn1B = Cross(rB0, normal);
            float rn2A = Cross(rA1, normal), rn2B = Cross(rB1, normal);
            float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
            float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
            float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
            const float k_maxConditionNumber = 1000.0f;
            if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
                // b2Mat22::GetInverse with ex = (k11, k12), ey = (k12, k22); K is symmetric, so is the inverse
                float det = k11 * k22 - k12 * k12;
                if (det != 0.0f) det = Div(1.0f, det);
                r5 = make_float4(normalMass, tangentMass, k11, k12);
                r6 = make_float4(k22, det * k22, -det * k12, det * k11);
            } else {
                vcPointCount = 1;
            }
        }
        float4 imp = make_float4(0, 0, 0, 0);
        if (warmStarting) imp = make_float4(dtRatio * p0.z, dtRatio * p0.w, dtRatio * p1.z, dtRatio * p1.w);
        float4* vr = velRec + (size_t)i * kVelRec * kTile;
        vr[0 * kTile] = make_float4(normal.x, normal.y, friction, __int_as_float(PackBodies(bodyA, bodyB, vcPointCount)));
        vr[1 * kTile] = make_float4(mA, iA, mB, iB);
        vr[2 * kTile] = r2;
        vr[3 * kTile] = r3;
        vr[4 * kTile] = r4;
        vr[5 * kTile] = r5;
        vr[6 * kTile] = r6;
        impRec[(size_t)i * kTile] = imp;
        float4* pr = posRec + (size_t)i * kPosRec * kTile;
        pr[0 * kTile] = nn;
        pr[1 * kTile] = make_float4(p0.x, p0.y, p1.x, p1.y);
        pr[2 * kTile] = make_float4(mA, iA, mB, iB);
        pr[3 * kTile] = make_float4(msA.z, msA.w, msB.z, msB.w);
        // manifold type | last contact of its island << 2 | island << 8
        pr[4 * kTile] = make_float4(radiusA, radiusB, __int_as_float(PackBodies(bodyA, bodyB, pointCount)),
                                    __int_as_float(mtype | (i == islEnd - 1 ? 4 : 0) | (isl << 8)));
        circles |= mtype == kManifoldCircles;
    }
    if (circles) atomicOr(&A.solveFlags[ws], 1);  // k_solve_position picks its e_circles-capable variant
}

// =================================================================================================
// k_solve_velocity: one tile per CTA, lane = world.  Warp 0 (consumer) runs WarmStart and the velIters
// Gauss-Seidel passes out of shared memory; warp 1 (producer) feeds it: per island contact one TMA bulk copy
// of the 7-float4 record of all 32 worlds plus a cp.async of the accumulated impulses into the same stage,
// both completing on the stage's "full" mbarrier, and writes the updated impulses back to global memory
// when the consumer hands the stage back through its "empty" mbarrier.
// =================================================================================================
constexpr int kVelStageBytes = (kVelRec + 1) * 16 * kTile;  // record + impulses

__device__ __forceinline__ uint32_t MbarTryWait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    // suspend-time hint: the producer warp sleeps in hardware until the phase completes instead of
    // spinning next to the consumer warp
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(SmemAddr(bar)), "r"(parity), "r"(20000u)
                 : "memory");
    return ok;
}
__device__ __forceinline__ void MbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(bar)) : "memory");
}
// arrive on `bar` once every cp.async this thread issued so far has landed (does not change the pending count)
__device__ __forceinline__ void CpAsyncArriveNoInc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(SmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void CpAsync16Cg(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(SmemAddr(dst)), "l"(src) : "memory");
}


// ---- shared-memory accessors on 32-bit shared-window addresses (no generic->shared conversion in the loops) ----
__device__ __forceinline__ float4 LdsF4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float LdsF(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t LdsU(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void StsU(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void StsF(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
__device__ __forceinline__ void StsF4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ uint32_t MbarTryWaitA(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}
__device__ __forceinline__ void MbarArriveA(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Opaque identity: stops the compiler from rematerialising a loop invariant inside the loop.
__device__ __forceinline__ uint32_t Keep(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ int Keep(int v) { asm volatile("" : "+r"(v)); return v; }
template <class P>
__device__ __forceinline__ P* Keep(P* v) { asm volatile("" : "+l"(v)); return v; }
// How the consumer learns that a ring stage has landed.  The stage is written by the ASYNC proxy (a TMA bulk copy, and in
// the velocity kernel 32 cp.async's that arrive on the same mbarrier); only a thread that itself observes the mbarrier's
// phase complete is guaranteed to see those bytes.  (Round 2's first build had a watcher lane observe the barriers and
// publish a count in a shared word the consumer polled: two runs of one 16,384-world batch then disagreed in a few tiles
// per thousand steps — the consumer now and then read a stage whose bytes were not all visible to it yet,
// tools/diag_determinism.py.)  So every consumer lane tests the barrier itself, but off the chain: the NON-blocking test
// for record k + 1 is issued while record k is being fetched, and its predicate is first looked at a whole contact solve
// later, when the 50 - 90 cycles an mbarrier test takes have long passed; only if that test said "not yet" does the warp
// fall into the blocking wait.  Stage s = k & (depth - 1) is reused every `depth` records and the look-ahead is shorter
// than that, so the parity of record k's use, (k / depth) & 1, cannot alias an older phase.
__device__ __forceinline__ uint32_t MbarTestA(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok;
}
__device__ __forceinline__ void MbarSpinA(uint32_t bar, uint32_t parity) {
    while (!MbarTryWaitA(bar, parity)) {}
}
__device__ __forceinline__ void CpAsync16CgA(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void CpAsyncArriveNoIncA(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// The tile's b2Velocity / b2Position table in shared memory: component c of body b of this lane's world.
struct BodyTable {
    uint32_t base;  // shared address of (body 0, component 0, this lane)
    __device__ __forceinline__ uint32_t at(int b) const { return base + (uint32_t)b * (3 * kTile * 4); }
};

// One velocity record in registers, with the shared-memory addresses of its two bodies' table rows.
struct VelRecord {
    float4 q0, q1, q2, q3, q4, q5, q6, imp;
    uint32_t pA, pB;  // table rows (a lane without a contact at this index points both at its scratch row)
    int points;       // manifold points of this lane's contact (0: no contact at this index, nothing to solve)
    int index;        // island contact index (where the accumulated impulses live in impRec)
};
// The six floats of a contact's two b2Velocity entries.
struct BodyVals {
    float ax, ay, aw, bx, by, bw;
};
__device__ __forceinline__ void IssueVelRecord(VelRecord& r, uint32_t st) {
    r.q0 = LdsF4(st + 0 * 512); r.q1 = LdsF4(st + 1 * 512); r.q2 = LdsF4(st + 2 * 512); r.q3 = LdsF4(st + 3 * 512);
    r.q4 = LdsF4(st + 4 * 512); r.q5 = LdsF4(st + 5 * 512); r.q6 = LdsF4(st + 6 * 512); r.imp = LdsF4(st + 7 * 512);
}
__device__ __forceinline__ void TouchF4(float4& v) { asm volatile("" : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)); }
// Decode the record issued by IssueVelRecord (this is the first use of the loaded registers, i.e. where the loads
// are waited for: it is placed a whole contact solve after IssueVelRecord).
__device__ __forceinline__ void FinishVelRecord(VelRecord& r, bool active, BodyTable T, uint32_t scratch) {
    TouchF4(r.q0); TouchF4(r.q1); TouchF4(r.q2); TouchF4(r.q3); TouchF4(r.q4); TouchF4(r.q5); TouchF4(r.q6); TouchF4(r.imp);
    const int packed = __float_as_int(r.q0.w);
    r.pA = active ? T.at(packed & 0x7fff) : scratch;
    r.pB = active ? T.at((packed >> 15) & 0x7fff) : scratch;
    const int pointCount = (packed >> 30) & 3;
    r.points = active ? (pointCount > 1 ? 2 : pointCount) : 0;
}
__device__ __forceinline__ BodyVals LoadBodyVals(uint32_t pA, uint32_t pB) {
    BodyVals b;
    b.ax = LdsF(pA); b.ay = LdsF(pA + 128); b.aw = LdsF(pA + 256);
    b.bx = LdsF(pB); b.by = LdsF(pB + 128); b.bw = LdsF(pB + 256);
    return b;
}
__device__ __forceinline__ void StoreBodyVals(uint32_t pA, uint32_t pB, const BodyVals& b) {
    StsF(pA, b.ax); StsF(pA + 128, b.ay); StsF(pA + 256, b.aw);
    StsF(pB, b.bx); StsF(pB + 128, b.by); StsF(pB + 256, b.bw);
}
// Store-to-load forwarding by hand: `loaded` was read from the table BEFORE `cur` (the values the previous contact
// left for its bodies, rows curA / curB) was stored, so rows the two contacts share are taken from `cur` instead.
__device__ __forceinline__ BodyVals ForwardBodyVals(const BodyVals& loaded, const BodyVals& cur, uint32_t pA, uint32_t pB,
                                                    uint32_t curA, uint32_t curB) {
    const bool aa = pA == curA, ab = pA == curB, ba = pB == curA, bb = pB == curB;
    BodyVals o;
    o.ax = aa ? cur.ax : (ab ? cur.bx : loaded.ax);
    o.ay = aa ? cur.ay : (ab ? cur.by : loaded.ay);
    o.aw = aa ? cur.aw : (ab ? cur.bw : loaded.aw);
    o.bx = ba ? cur.ax : (bb ? cur.bx : loaded.bx);
    o.by = ba ? cur.ay : (bb ? cur.by : loaded.by);
    o.bw = ba ? cur.aw : (bb ? cur.bw : loaded.bw);
    return o;
}

// WarmStart of one contact (b2ContactSolver.cpp:240-278) on body values held in registers
__device__ __forceinline__ void WarmStartContact(const VelRecord& r, BodyVals& bv) {
    const int points = r.points;
    const float mA = r.q1.x, iA = r.q1.y, mB = r.q1.z, iB = r.q1.w;
    Vec2 vA = V2(bv.ax, bv.ay), vB = V2(bv.bx, bv.by);
    float wA = bv.aw, wB = bv.bw;
    const Vec2 normal = V2(r.q0.x, r.q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    if (points > 0) {
        Vec2 P = r.imp.x * normal + r.imp.y * tangent;
        wA -= iA * Cross(V2(r.q2.x, r.q2.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(r.q2.z, r.q2.w), P);
        vB = vB + mB * P;
    }
    if (points > 1) {
        Vec2 P = r.imp.z * normal + r.imp.w * tangent;
        wA -= iA * Cross(V2(r.q4.x, r.q4.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(r.q4.z, r.q4.w), P);
        vB = vB + mB * P;
    }
    bv.ax = vA.x; bv.ay = vA.y; bv.aw = wA;
    bv.bx = vB.x; bv.by = vB.y; bv.bw = wB;
}

// One point's friction constraint (b2ContactSolver.cpp:305-333; tangentSpeed is 0: b2Contact::SetTangentSpeed is
// not exposed)
__device__ __forceinline__ void SolveTangent(Vec2& vA, float& wA, Vec2& vB, float& wB, Vec2 rA, Vec2 rB, Vec2 tangent,
                                             float tangentMass, float friction, float normalImpulse, float& tangentImpulse,
                                             float mA, float iA, float mB, float iB) {
    Vec2 dv = vB + Cross(wB, rB) - vA - Cross(wA, rA);
    float vt = Dot(dv, tangent) - 0.0f;
    float lambda = tangentMass * (-vt);
    float maxFriction = friction * normalImpulse;
    float newImpulse = Clamp(tangentImpulse + lambda, -maxFriction, maxFriction);
    lambda = newImpulse - tangentImpulse;
    tangentImpulse = newImpulse;
    Vec2 P = lambda * tangent;
    vA = vA - mA * P;
    wA -= iA * Cross(rA, P);
    vB = vB + mB * P;
    wB += iB * Cross(rB, P);
}

// SolveVelocityConstraints for one contact (b2ContactSolver.cpp:280-597), straight-line for every point count: the
// friction constraint of point 0 is common; then the one-point normal constraint and (friction of point 1 + block
// solver) are both evaluated — they are independent, the one-point arithmetic fills issue slots the two-point chain
// leaves empty — and the result of the contact's own point count is selected.  The reference tries the block
// solver's four cases in order and takes the first feasible one; here all four are evaluated and the same one is
// selected.  `points` == 0 (no contact for this lane at this index): any result will do, it goes to the scratch row.
__device__ __forceinline__ float4 SolveVelocityContact(const VelRecord& r, BodyVals& bv) {
    const float4 q0 = r.q0, q1 = r.q1, q2 = r.q2, q3 = r.q3, q4 = r.q4, q5 = r.q5, q6 = r.q6;
    float4 imp = r.imp;
    const float mA = q1.x, iA = q1.y, mB = q1.z, iB = q1.w;
    Vec2 vA = V2(bv.ax, bv.ay), vB = V2(bv.bx, bv.by);
    float wA = bv.aw, wB = bv.bw;
    const Vec2 normal = V2(q0.x, q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    const Vec2 rA0 = V2(q2.x, q2.y), rB0 = V2(q2.z, q2.w);
    const Vec2 rA1 = V2(q4.x, q4.y), rB1 = V2(q4.z, q4.w);
    const float friction = q0.z;
    SolveTangent(vA, wA, vB, wB, rA0, rB0, tangent, q3.y, friction, imp.x, imp.y, mA, iA, mB, iB);
    // ---- one point: normal constraint (b2ContactSolver.cpp:336-371) ----
    Vec2 vA1 = vA, vB1 = vB;
    float wA1 = wA, wB1 = wB, impX1;
    {
        Vec2 dv = vB + Cross(wB, rB0) - vA - Cross(wA, rA0);
        float vn = Dot(dv, normal);
        float lambda = -q3.x * (vn - q3.z);
        float newImpulse = Max(imp.x + lambda, 0.0f);
        lambda = newImpulse - imp.x;
        impX1 = newImpulse;
        Vec2 P = lambda * normal;
        vA1 = vA1 - mA * P;
        wA1 -= iA * Cross(rA0, P);
        vB1 = vB1 + mB * P;
        wB1 += iB * Cross(rB0, P);
    }
    // ---- two points: friction of point 1, then the block solver (b2ContactSolver.cpp:388-589) ----
    SolveTangent(vA, wA, vB, wB, rA1, rB1, tangent, q5.y, friction, imp.z, imp.w, mA, iA, mB, iB);
    const float k11 = q5.z, k12 = q5.w, k22 = q6.x;
    const float nmxx = q6.y, nmxy = q6.z, nmyy = q6.w;  // normalMass = K^-1 (symmetric)
    const Vec2 a = V2(imp.x, imp.z);
    Vec2 dv1 = vB + Cross(wB, rB0) - vA - Cross(wA, rA0);
    Vec2 dv2 = vB + Cross(wB, rB1) - vA - Cross(wA, rA1);
    float vn1 = Dot(dv1, normal);
    float vn2 = Dot(dv2, normal);
    Vec2 b;
    b.x = vn1 - q3.z;
    b.y = vn2 - q3.w;
    b.x -= k11 * a.x + k12 * a.y;
    b.y -= k12 * a.x + k22 * a.y;
    const float x1 = -(nmxx * b.x + nmxy * b.y);  // case 1: x = -normalMass * b
    const float y1 = -(nmxy * b.x + nmyy * b.y);
    const bool ok1 = x1 >= 0.0f && y1 >= 0.0f;
    const float x2 = -q3.x * b.x;                  // case 2: x2 = 0
    const bool ok2 = x2 >= 0.0f && (k12 * x2 + b.y) >= 0.0f;
    const float y3 = -q5.x * b.y;                  // case 3: x1 = 0
    const bool ok3 = y3 >= 0.0f && (k12 * y3 + b.x) >= 0.0f;
    const bool ok4 = b.x >= 0.0f && b.y >= 0.0f;   // case 4: both 0
    const bool solved = ok1 || ok2 || ok3 || ok4;
    Vec2 x;
    x.x = ok1 ? x1 : (ok2 ? x2 : 0.0f);
    x.y = ok1 ? y1 : ((!ok2 && ok3) ? y3 : 0.0f);
    Vec2 d = x - a;
    Vec2 P1 = d.x * normal;
    Vec2 P2 = d.y * normal;
    Vec2 nvA = vA - mA * (P1 + P2);
    float nwA = wA - iA * (Cross(rA0, P1) + Cross(rA1, P2));
    Vec2 nvB = vB + mB * (P1 + P2);
    float nwB = wB + iB * (Cross(rB0, P1) + Cross(rB1, P2));
    // ---- select: one point | two points solved | two points, no feasible case (velocities and impulses stay) ----
    const bool onePoint = r.points == 1;
    const bool take = solved && !onePoint;
    bv.ax = onePoint ? vA1.x : (take ? nvA.x : vA.x);
    bv.ay = onePoint ? vA1.y : (take ? nvA.y : vA.y);
    bv.aw = onePoint ? wA1 : (take ? nwA : wA);
    bv.bx = onePoint ? vB1.x : (take ? nvB.x : vB.x);
    bv.by = onePoint ? vB1.y : (take ? nvB.y : vB.y);
    bv.bw = onePoint ? wB1 : (take ? nwB : wB);
    imp.x = onePoint ? impX1 : (take ? x.x : imp.x);
    imp.z = onePoint ? r.imp.z : (take ? x.y : imp.z);
    imp.w = onePoint ? r.imp.w : imp.w;
    return imp;
}

// JOINTS: the batch has joints (a separate instantiation, so that the joint-free kernel's code is what it was).
template <bool JOINTS>
__global__ void __launch_bounds__(128) k_solve_velocity(Arena A, int velIters, float h) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const int n = A.solveCount[ws];
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    // joints of this world's islands, in island order (k_islands): b2Island::Solve initialises them after the contacts'
    // warm start and solves them in front of the contacts in every velocity iteration (b2Island.cpp:252-269).  Islands
    // only share static bodies, so "every island's joints, then every island's contacts" is the same arithmetic.
    int nJoints = 0;
    if (JOINTS) {
        const int islands = A.islandCount[ws];
        if (islands > 0 && A.jointCount[ws] > 0) nJoints = col(A.islandJointEnd, tile, A.NB, lane)[islands - 1];
    }
    const bool anyJoints = JOINTS && __any_sync(0xffffffffu, nJoints > 0);
    if (nMax == 0 && !anyJoints) return;  // nothing to solve in the whole tile: velocities stay as integrated
    const bool warm = (A.worldFlags[ws] & B2G_WORLD_WARM_STARTING) != 0;
    const int anyWarm = __any_sync(0xffffffffu, warm) ? 1 : 0;
    const int total = nMax * (velIters + anyWarm);
    // Two ring protocols.
    //  * pipelined (nMax >= kVelPipelineMin, 8 stages): a record is loaded into registers two steps before it is solved
    //    and its stage is released once it HAS been solved (B2G_EARLY_RELEASE explains why not sooner); the consumer
    //    writes a contact's updated impulses straight to impRec, and the producer fetches them again for the next pass
    //    a whole pass (>= kVelPipelineMin - 8 steps, and at least one release / acquire pair on an `empty` barrier)
    //    later.
    //  * small worlds: one record at a time; the updated impulses ride back in the ring stage and the producer lane that
    //    fetched them writes them to impRec when the stage is recycled — store and re-fetch of one contact are issued by
    //    the same thread, in that order.
    // The depth is a power of two (stage of record k = k & (depth - 1)).
    // A tile with joints takes the one-record path: the joints run between two passes on the table in shared memory,
    // and the pipelined consumer carries the next contact's body values in registers across pass boundaries.
    const bool pipelined = nMax >= kVelPipelineMin && !anyJoints;
    int depth = 1, depthLog = 0;
    while (depth * 2 <= nMax - 1 && depth * 2 <= kVelStages) { depth *= 2; ++depthLog; }

    const uint32_t bodyBytes = (uint32_t)A.NB * 3 * kTile * 4;
    unsigned char* stages = smem + SolveTableBytes(A.NB);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + kVelStages * kVelStageBytes);
    uint64_t* empty = full + kVelStages;
    uint64_t* bodyBar = empty + kVelStages;
    float* gVel = A.solveVel + (size_t)tile * A.NB * 3 * kTile;
    float4* gImp = A.impRec + (size_t)tile * A.NC * kTile + lane;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kVelStages; ++s) {
            MbarInit(&full[s], 33);   // expect_tx arrive of the TMA issuer + 32 cp.async completions
            MbarInit(&empty[s], 1);   // one consumer lane, on behalf of the (converged) warp
        }
        MbarInit(bodyBar, 1);
        FenceBarrierInit();
    }
    __syncthreads();

    const int role = SequentialRole(warp);
    if (role != kRoleConsumer && role != kRoleProducer) return;
    if (role == kRoleProducer) {
        // ================= producer =================
        if (lane == 0) {
            MbarExpectTx(bodyBar, bodyBytes);
            TmaLoad1D(smem, gVel, bodyBytes, bodyBar);
        }
        const unsigned char* src = reinterpret_cast<const unsigned char*>(A.velRec + (size_t)tile * A.NC * kVelRec * kTile);
        const bool fence = nMax - 1 <= depth;  // tiny worlds: write-back and re-fetch of a contact are back to back
        int i = 0, s = 0;
        int wi = 0, wpass = 0;  // record that previously occupied stage s (write-back cursor)
        uint32_t phase = 1;     // parity of the consumer's release that frees stage s for its next use
        // (experiment, off by default: records pulled into L2 kVelPrefetch records ahead of the ring's own loads)
        int pi = 0;  // record index of the next prefetch, pf records issued so far
        int pf = 0;
        auto prefetch = [&]() {
            if (lane == 0 && pf < total && kVelPrefetch > 0) TmaPrefetchL2(src + (size_t)pi * (kVelRec * 16 * kTile), kVelRec * 16 * kTile);
            ++pf;
            if (++pi == nMax) pi = 0;
        };
        for (int k = 0; k < kVelPrefetch; ++k) prefetch();
        for (int seq = 0; seq < total + depth; ++seq) {
            float4* st = reinterpret_cast<float4*>(stages + (size_t)s * kVelStageBytes) + lane;
            if (seq >= depth) {  // stage s holds record seq - depth: wait for the consumer (small worlds: write its impulses back)
                if (pipelined && seq >= total) break;  // nothing left to load, nothing to write back
                while (!MbarTryWait(&empty[s], phase)) {}
                if (!pipelined) {
                    if (wi < n && !(anyWarm && wpass == 0)) gImp[(size_t)wi * kTile] = st[7 * kTile];
                    if (fence) __threadfence();
                    if (++wi == nMax) { wi = 0; ++wpass; }
                }
            }
            if (seq < total) {
                prefetch();
                if (lane == 0) {
                    MbarExpectTx(&full[s], kVelRec * 16 * kTile);
                    TmaLoad1D(st, src + (size_t)i * (kVelRec * 16 * kTile), kVelRec * 16 * kTile, &full[s]);
                }
                if (i < n) CpAsync16Cg(&st[7 * kTile], gImp + (size_t)i * kTile);
                CpAsyncArriveNoInc(&full[s]);
                if (++i == nMax) i = 0;
            }
            if (++s == depth) { s = 0; phase ^= 1u; }
        }
        // pipelined: the loop ends with the last records' impulse copies still in flight; a thread must not exit on them
        asm volatile("cp.async.wait_all;" ::: "memory");
        return;
    }

    // ================= consumer =================
    // One world's chain is ~velIters * n dependent contact solves, so everything that is not the solve's own arithmetic
    // is kept off it:
    //  * three records live in registers (being solved / next / being fetched).  A record's shared-memory loads are
    //    issued at the top of a step and first looked at at the end of it, a whole solve later;
    //  * the body values of the NEXT contact are read from the table at the top of the step too; rows both contacts
    //    touch are forwarded from registers at the end (ForwardBodyVals) instead of going through a store and a reload;
    //  * lanes without a contact at the current index solve a scratch row instead of branching around the solve, and
    //    the solve itself is straight-line for every point count (SolveVelocityContact): a step is one basic block;
    //  * whether the next record has landed is asked one step ahead (MbarTestA), so the answer costs the chain nothing.
    while (!MbarTryWait(bodyBar, 0)) {}
    // Loop invariants are pinned in registers (Keep): left alone, the compiler re-derives them from %tid / %ctaid /
    // the shared-window base inside the loop, which costs more than the solve's own address arithmetic.
    const BodyTable T{Keep(SmemAddr(smem) + (uint32_t)lane * 4)};
    const uint32_t scratch = Keep(T.at(A.NB));  // row NB of the table: this lane's scratch row
    const uint32_t stage0 = Keep(SmemAddr(stages) + (uint32_t)lane * 16);
    const uint32_t empty0 = Keep(SmemAddr(empty));
    const uint32_t full0 = Keep(SmemAddr(full));
    const uint32_t mask = Keep((uint32_t)depth - 1u);
    const uint32_t depthShift = Keep((uint32_t)depthLog);
    const int warmRecords = Keep(anyWarm ? nMax : 0);
    const int nK = Keep(n), nMaxK = Keep(nMax), totalK = Keep(total);

    if (pipelined) {
        int iLoad = 0, seq = 0;  // seq: record being solved; records up to seq + 1 are in registers
        auto stageOf = [&](int k) { return stage0 + (((uint32_t)k & mask) << 12); };  // kVelStageBytes == 4096
        auto emptyOf = [&](int k) { return empty0 + (((uint32_t)k & mask) << 3); };
        auto fullOf = [&](int k) { return full0 + (((uint32_t)k & mask) << 3); };
        auto parityOf = [&](int k) { return ((uint32_t)k >> depthShift) & 1u; };
        uint32_t landedNext;  // the test of record seq + 2's barrier, issued a step ago
        // record k (issued earlier) becomes usable
        auto finish = [&](VelRecord& r, int k) {
            FinishVelRecord(r, k < totalK && iLoad < nK, T, scratch);
            r.index = iLoad;
            if (++iLoad == nMaxK) iLoad = 0;
            __syncwarp();
            if (B2G_EARLY_RELEASE && k < totalK && lane == 0) MbarArriveA(emptyOf(k));
        };
        // One chain step: `cur` (record seq) is solved on `bv`; meanwhile `fill` receives record seq + 2 (past the last
        // record: a stale stage, read and ignored) and the body values of `nxt` are read.
        auto step = [&](VelRecord& cur, const VelRecord& nxt, VelRecord& fill, BodyVals& bv, auto warmPass) {
            constexpr bool WARM = decltype(warmPass)::value;
            if (!landedNext && seq + 2 < totalK) MbarSpinA(fullOf(seq + 2), parityOf(seq + 2));
            IssueVelRecord(fill, stageOf(seq + 2));
            landedNext = MbarTestA(fullOf(seq + 3), parityOf(seq + 3));  // past the last record: an answer nobody uses
            const BodyVals nextLoaded = LoadBodyVals(nxt.pA, nxt.pB);  // before this contact's stores: ForwardBodyVals
            if (WARM) {
                if (warm) WarmStartContact(cur, bv);
            } else {
                const float4 imp = SolveVelocityContact(cur, bv);
                if (cur.points > 0) gImp[(size_t)cur.index * kTile] = imp;
            }
            StoreBodyVals(cur.pA, cur.pB, bv);
            bv = ForwardBodyVals(nextLoaded, bv, nxt.pA, nxt.pB, cur.pA, cur.pB);
            finish(fill, seq + 2);
            // record seq is solved and its results are stored: its stage goes back to the producer (see B2G_EARLY_RELEASE)
            if (!B2G_EARLY_RELEASE && lane == 0) MbarArriveA(emptyOf(seq));
            ++seq;
        };
        VelRecord r0, r1, r2;
        MbarSpinA(fullOf(0), parityOf(0));
        IssueVelRecord(r0, stageOf(0));
        finish(r0, 0);
        MbarSpinA(fullOf(1), parityOf(1));
        IssueVelRecord(r1, stageOf(1));
        finish(r1, 1);
        landedNext = MbarTestA(fullOf(2), parityOf(2));
        BodyVals bv = LoadBodyVals(r0.pA, r0.pB);
        // warm-start pass, and whatever does not fill a group of three: one step at a time, registers rotated by moves
        auto single = [&](auto warmPass) {
            step(r0, r1, r2, bv, warmPass);
            r0 = r1;
            r1 = r2;
        };
        while (seq < warmRecords) single(std::true_type());
        while (seq + 3 <= totalK) {
            step(r0, r1, r2, bv, std::false_type());
            step(r1, r2, r0, bv, std::false_type());
            step(r2, r0, r1, bv, std::false_type());
        }
        while (seq < totalK) single(std::false_type());
    } else {
        const JointCols jc = MakeJointCols(A, tile, lane, T.base, 0);
        Col<int> iJoints = col(A.islandJoints, tile, A.NJ, lane);
        const float dtRatio = A.invDt0[ws] * h;
        auto jointsInit = [&]() {
            for (int j = 0; j < nJoints; ++j) JointInitVelocity(&jc, iJoints[j], h, dtRatio, warm);
            __syncwarp();
        };
        auto jointsSolve = [&]() {
            for (int j = 0; j < nJoints; ++j) JointSolveVelocity(&jc, iJoints[j], h);
            __syncwarp();
        };
        int i = 0, cs = 0;
        for (int seq = 0; seq < total; ++seq) {
            if (anyJoints && i == 0 && seq >= warmRecords) {  // a velocity iteration starts
                if (seq == warmRecords) jointsInit();
                jointsSolve();
            }
            const uint32_t st = stage0 + cs * kVelStageBytes;
            MbarSpinA(full0 + cs * 8, ((uint32_t)seq >> depthShift) & 1u);
            if (i < n) {
                VelRecord r;
                IssueVelRecord(r, st);
                FinishVelRecord(r, true, T, scratch);
                BodyVals bv = LoadBodyVals(r.pA, r.pB);
                if (seq < warmRecords) {
                    if (warm) WarmStartContact(r, bv);
                } else {
                    StsF4(st + 7 * 512, SolveVelocityContact(r, bv));
                }
                StoreBodyVals(r.pA, r.pB, bv);
            }
            __syncwarp();
            if (lane == 0) MbarArriveA(empty0 + cs * 8);
            if (++i == nMax) i = 0;
            if (++cs == depth) cs = 0;
        }
        if (anyJoints && (nMax == 0 || velIters == 0)) {  // no contact pass to hang the joints on
            jointsInit();
            for (int it = 0; it < velIters; ++it) jointsSolve();
        }
    }
    // velocities back to the solver table in global memory
    FenceProxyAsync();
    __syncwarp();
    if (lane == 0) {
        TmaStore1D(gVel, smem, bodyBytes);
        TmaStoreCommitAndWait();
    }
}

// =================================================================================================
// k_solve_integrate: StoreImpulses, PostSolve events, integrate positions -> position table
// =================================================================================================
__global__ void __launch_bounds__(kPrepWarps * 32) k_solve_integrate(Arena A, float h) {
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const bool live = tile * kTile + lane < A.numWorlds;
    const int nb = live ? A.bodyCount[ws] : 0;
    const int n = live ? A.solveCount[ws] : 0;

    // ---- StoreImpulses + Report ----
    Col<int> iContacts = col(A.islandContacts, tile, A.NC, lane);
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane);
    Col<float4> cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    const float4* velRec = A.velRec + (size_t)tile * A.NC * kVelRec * kTile + lane;
    const float4* impRec = A.impRec + (size_t)tile * A.NC * kTile + lane;
    const bool post = (A.eventMask >> kEventPostSolve & 1) != 0;
    const int evBase = post ? A.eventCount[ws] : 0;  // events of this step so far (Collide)
    Col<int4> evA = col(A.eventA, tile, A.NE, lane);
    Col<float4> evB = col(A.eventB, tile, A.NE, lane);
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    for (int i = warp; i < nMax; i += kPrepWarps) {
        if (i >= n) continue;
        const int k = iContacts[i];
        const float4 imp = impRec[(size_t)i * kTile];
        const int vcPointCount = (__float_as_int(velRec[(size_t)i * kVelRec * kTile].w) >> 30) & 3;
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
        if (post) {  // b2Island::Report: islands in order, contacts in island order == i order
            const int slot = evBase + i;
            if (slot < A.NE) {
                const int4 a = cA[k];
                int pa = proxyInfo[a.x & 0xffff], pb = proxyInfo[(a.x >> 16) & 0xffff];
                evA[slot] = make_int4(kEventPostSolve, pa, pb, vcPointCount);
                evB[slot] = make_float4(imp.x, vcPointCount > 1 ? imp.z : 0.0f, imp.y, vcPointCount > 1 ? imp.w : 0.0f);
            }
        }
    }
    if (post && warp == 0 && n > 0) {
        if (evBase + n > A.NE) A.status[ws] |= kStatusEventOverflow;
        A.eventCount[ws] = evBase + n;
    }

    // ---- integrate positions (b2Island.cpp:276-306) ----
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    const float* sv = A.solveVel + (size_t)tile * A.NB * 3 * kTile + lane;
    float* sp = A.solvePos + (size_t)tile * A.NB * 3 * kTile + lane;
    const int nbMax = __reduce_max_sync(0xffffffffu, nb);
    for (int b = warp; b < nbMax; b += kPrepWarps) {
        if (b >= nb) continue;
        const float4 sw = bodySweep[b];
        Vec2 c = V2(sw.x, sw.y);
        float a = sw.z;
        if (bodyFlags[b] & kBodyIsland) {
            Vec2 v = V2(sv[(size_t)(b * 3 + 0) * kTile], sv[(size_t)(b * 3 + 1) * kTile]);
            float wv = sv[(size_t)(b * 3 + 2) * kTile];
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
            bodyVel[b] = make_float4(v.x, v.y, wv, 0.0f);
        }
        sp[(size_t)(b * 3 + 0) * kTile] = c.x;
        sp[(size_t)(b * 3 + 1) * kTile] = c.y;
        sp[(size_t)(b * 3 + 2) * kTile] = a;
    }
}

// =================================================================================================
// k_solve_position: one tile per CTA, lane = world.  Up to posIters NGS passes (b2Island.cpp:308-328,
// b2ContactSolver.cpp:614-744); each island stops as soon as one of its passes ends with
// minSeparation >= -3 * b2_linearSlop.  Warp 0 (consumer) solves out of shared memory, lane 0 of warp 1
// (producer) streams the position records of one pass through a TMA + mbarrier ring and then waits for the
// consumer's verdict on whether another pass is needed.  The consumer is built like k_solve_velocity's: three records in registers, the next contact's
// body values fetched during the solve and forwarded, lanes without work on a scratch row, a step is
// straight-line code.
// =================================================================================================
constexpr int kPosStageBytes = kPosRec * 16 * kTile;  // 2560

struct PosRecord {
    float4 r0, r1, r2, r3, r4;
    uint32_t pA, pB;  // table rows of the two bodies
    int info;         // manifold type | last contact of its island << 2 | has a contact << 3 | two points << 4 | island << 8
};
__device__ __forceinline__ void IssuePosRecord(PosRecord& r, uint32_t st) {
    r.r0 = LdsF4(st); r.r1 = LdsF4(st + 512); r.r2 = LdsF4(st + 1024); r.r3 = LdsF4(st + 1536); r.r4 = LdsF4(st + 2048);
}
__device__ __forceinline__ void FinishPosRecord(PosRecord& r, bool active, BodyTable T) {
    TouchF4(r.r0); TouchF4(r.r1); TouchF4(r.r2); TouchF4(r.r3); TouchF4(r.r4);
    const int packed = __float_as_int(r.r4.z);
    r.pA = T.at(packed & 0x7fff);
    r.pB = T.at((packed >> 15) & 0x7fff);
    const int pointCount = (packed >> 30) & 3;
    const int word = __float_as_int(r.r4.w);
    r.info = active && pointCount > 0 ? ((word & ~0x18) | 8 | (pointCount > 1 ? 16 : 0)) : 0;
}

// One manifold point of SolvePositionConstraints (b2ContactSolver.cpp:687-734; b2PositionSolverManifold :614-664).
// e_faceA / e_faceB are the same arithmetic with the bodies' roles swapped: selected, not branched on (lane = world,
// a warp holds both kinds); CIRCLES adds the e_circles arithmetic beside it, again selected.
// ROT: how sinf / cosf are evaluated.  kRotNarrow: the |angle| < pi/4 arithmetic only; kRotWide: the |angle| < 120
// arithmetic (quadrant reduction, still branch-free: tiles with bodies that roll or tumble); `slow` is set when that was
// not valid for one of the two angles (the caller then redoes the contact with kRotAny, the full function).
enum : int { kRotAny = 0, kRotNarrow = 1, kRotWide = 2 };
template <bool CIRCLES, int ROT>
__device__ __forceinline__ float SolvePositionPoint(const PosRecord& r, Vec2 lp, Vec2& cAp, float& aA, Vec2& cBp, float& aB,
                                                    bool& slow) {
    const int mtype = r.info & 3;
    const float mA = r.r2.x, iA = r.r2.y, mB = r.r2.z, iB = r.r2.w;
    const Vec2 localCenterA = V2(r.r3.x, r.r3.y), localCenterB = V2(r.r3.z, r.r3.w);
    const Vec2 localNormal = V2(r.r0.x, r.r0.y), localPoint = V2(r.r0.z, r.r0.w);
    const float radiusA = r.r4.x, radiusB = r.r4.y;
    Xf xfA, xfB;
    if (ROT == kRotNarrow) {
        slow = slow || !InSinCosFastRange(aA) || !InSinCosFastRange(aB);
        xfA.q = MakeRotFastRange(aA);
        xfB.q = MakeRotFastRange(aB);
    } else if (ROT == kRotWide) {
        slow = slow || !InSinCosWideRange(aA) || !InSinCosWideRange(aB);
        xfA.q = MakeRotWideRange(aA);
        xfB.q = MakeRotWideRange(aB);
    } else {
        xfA.q = MakeRot(aA);
        xfB.q = MakeRot(aB);
    }
    xfA.p = cAp - Mul(xfA.q, localCenterA);
    xfB.p = cBp - Mul(xfB.q, localCenterB);
    const bool faceA = mtype == kManifoldFaceA;
    Xf xfR, xfI;  // the reference face's body, the incident one
    xfR.q.s = faceA ? xfA.q.s : xfB.q.s; xfR.q.c = faceA ? xfA.q.c : xfB.q.c;
    xfR.p.x = faceA ? xfA.p.x : xfB.p.x; xfR.p.y = faceA ? xfA.p.y : xfB.p.y;
    xfI.q.s = faceA ? xfB.q.s : xfA.q.s; xfI.q.c = faceA ? xfB.q.c : xfA.q.c;
    xfI.p.x = faceA ? xfB.p.x : xfA.p.x; xfI.p.y = faceA ? xfB.p.y : xfA.p.y;
    Vec2 normal = Mul(xfR.q, localNormal);
    const Vec2 planePoint = Mul(xfR, localPoint);
    const Vec2 clipPoint = Mul(xfI, lp);
    float separation = Dot(clipPoint - planePoint, normal) - radiusA - radiusB;
    Vec2 point = clipPoint;
    normal.x = faceA ? normal.x : -normal.x;
    normal.y = faceA ? normal.y : -normal.y;
    if (CIRCLES) {
        const Vec2 pointA = Mul(xfA, localPoint);
        const Vec2 pointB = Mul(xfB, lp);
        Vec2 n = pointB - pointA;
        Normalize(n);
        const float sep = Dot(pointB - pointA, n) - radiusA - radiusB;
        const Vec2 mid = 0.5f * (pointA + pointB);
        const bool circ = mtype == kManifoldCircles;
        normal.x = circ ? n.x : normal.x; normal.y = circ ? n.y : normal.y;
        point.x = circ ? mid.x : point.x; point.y = circ ? mid.y : point.y;
        separation = circ ? sep : separation;
    }
    const Vec2 rA = point - cAp;
    const Vec2 rB = point - cBp;
    const float C = Clamp(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
    const float rnA = Cross(rA, normal);
    const float rnB = Cross(rB, normal);
    const float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
    const float quotient = Div(-C, K);  // unconditionally: a select below, not a branch around the division
    const float impulse = K > 0.0f ? quotient : 0.0f;
    const Vec2 P = impulse * normal;
    cAp = cAp - mA * P;
    aA -= iA * Cross(rA, P);
    cBp = cBp + mB * P;
    aB += iB * Cross(rB, P);
    return separation;
}

// A whole contact through the full-range sinf / cosf and every manifold type: out of line, for the rare contact one of
// whose bodies is turned further than pi/4.  Returns the bodies' new positions and min(separation) over its points.
struct PosContactResult {
    BodyVals bv;
    float minSeparation;
};
__device__ __noinline__ PosContactResult SolvePositionContactAnyAngle(float4 r0, float4 r1, float4 r2, float4 r3, float4 r4,
                                                                         int info, BodyVals bv) {
    PosRecord r;
    r.r0 = r0; r.r1 = r1; r.r2 = r2; r.r3 = r3; r.r4 = r4;
    r.pA = r.pB = 0;
    r.info = info;
    Vec2 cAp = V2(bv.ax, bv.ay), cBp = V2(bv.bx, bv.by);
    float aA = bv.aw, aB = bv.bw;
    bool slow = false;
    PosContactResult o;
    o.minSeparation = SolvePositionPoint<true, kRotAny>(r, V2(r1.x, r1.y), cAp, aA, cBp, aB, slow);
    if (info & 16) o.minSeparation = Min(o.minSeparation, SolvePositionPoint<true, kRotAny>(r, V2(r1.z, r1.w), cAp, aA, cBp, aB, slow));
    o.bv.ax = cAp.x; o.bv.ay = cAp.y; o.bv.aw = aA;
    o.bv.bx = cBp.x; o.bv.by = cBp.y; o.bv.bw = aB;
    return o;
}

template <bool JOINTS>
__global__ void __launch_bounds__(128) k_solve_position(Arena A, int posIters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int sContinue;
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const int n = A.solveCount[ws];
    const int islands = A.islandCount[ws];
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    const int doneWords = (A.NB + 31) / 32;
    int* gDone = A.islandSolved + (size_t)tile * doneWords * kTile + lane;  // bit per island: position solver converged
    Col<int> iContactEnd = col(A.islandContactEnd, tile, A.NB, lane);
    // Joints (b2Island.cpp:316-320): every position iteration of an island solves its contacts, then its joints, and the
    // island has converged once both say so.  A world with joints therefore gets its islands' verdicts after the joints
    // have run, behind the tile's contact pass (islands only share static bodies: the order across islands is free).
    Col<int> iJointEnd = col(A.islandJointEnd, tile, A.NB, lane);
    Col<int> iJoints = col(A.islandJoints, tile, A.NJ, lane);
    int nJoints = 0;
    if (JOINTS && islands > 0 && A.jointCount[ws] > 0) nJoints = iJointEnd[islands - 1];
    const bool anyJoints = JOINTS && __any_sync(0xffffffffu, nJoints > 0);
    if ((nMax == 0 && !anyJoints) || posIters <= 0) {
        // islands without contacts converge in the first pass (minSeparation == 0); no pass, no convergence
        if (warp == 0)
            for (int wd = 0; wd < doneWords; ++wd) gDone[(size_t)wd * kTile] = posIters > 0 ? -1 : 0;
        return;
    }
    const bool anyCircles = __any_sync(0xffffffffu, (A.solveFlags[ws] & 1) != 0);

    const uint32_t bodyBytes = (uint32_t)A.NB * 3 * kTile * 4;
    unsigned char* stages = smem + SolveTableBytes(A.NB);
    uint32_t* sDone = reinterpret_cast<uint32_t*>(stages + kRecStages * kPosStageBytes) + lane;  // word wd at sDone[wd * 32]
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + kRecStages * kPosStageBytes + (size_t)2 * doneWords * kTile * 4);
    uint64_t* empty = full + kRecStages;
    uint64_t* bodyBar = empty + kRecStages;
    uint64_t* passBar = bodyBar + 1;
    float* gPos = A.solvePos + (size_t)tile * A.NB * 3 * kTile;
    int depth = 1, depthLog = 0;  // a power of two: stage of record k = k & (depth - 1), k counted over all passes
    while (depth * 2 <= nMax && depth * 2 <= kRecStages) { depth *= 2; ++depthLog; }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRecStages; ++s) {
            MbarInit(&full[s], 1);    // the TMA issuer's expect_tx arrive
            MbarInit(&empty[s], 1);   // one consumer lane, on behalf of the (converged) warp
        }
        MbarInit(bodyBar, 1);
        MbarInit(passBar, 32);
        FenceBarrierInit();
    }
    __syncthreads();

    const int role = SequentialRole(warp);
    if (role != kRoleConsumer && role != kRoleProducer) return;
    if (role == kRoleProducer) {
        // ================= producer (one lane) =================
        if (lane != 0) return;
        MbarExpectTx(bodyBar, bodyBytes);
        TmaLoad1D(smem, gPos, bodyBytes, bodyBar);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(A.posRec + (size_t)tile * A.NC * kPosRec * kTile);
        int k = 0;  // records issued so far
        for (int pass = 0; pass < posIters; ++pass) {
            for (int i = 0; i < kPosPrefetch && i < nMax; ++i) TmaPrefetchL2(src + (size_t)i * kPosStageBytes, kPosStageBytes);  // off
            for (int i = 0; i < nMax; ++i, ++k) {
                if (kPosPrefetch > 0 && i + kPosPrefetch < nMax) TmaPrefetchL2(src + (size_t)(i + kPosPrefetch) * kPosStageBytes, kPosStageBytes);
                const int s = k & (depth - 1), use = k / depth;
                if (use > 0) while (!MbarTryWait(&empty[s], (uint32_t)(use - 1) & 1u)) {}
                MbarExpectTx(&full[s], kPosStageBytes);
                TmaLoad1D(stages + (size_t)s * kPosStageBytes, src + (size_t)i * kPosStageBytes, kPosStageBytes, &full[s]);
            }
            while (!MbarTryWait(passBar, (uint32_t)pass & 1u)) {}
            if (*(volatile int*)&sContinue == 0) break;
        }
        return;
    }

    // ================= consumer =================
    for (int wd = 0; wd < (JOINTS ? 2 : 1) * doneWords; ++wd) sDone[wd * kTile] = 0u;  // words doneWords..: "contacts fine this pass"
    // islands that own at least one contact (or joint) and have not converged yet
    const bool deferred = JOINTS && nJoints > 0;  // this world's islands get their verdict after the joints
    int pending = 0;
    {
        int prev = 0, prevJ = 0;
        for (int isl = 0; isl < islands; ++isl) {
            int e = iContactEnd[isl];
            const int ej = deferred ? iJointEnd[isl] : 0;
            if (e > prev || ej > prevJ) ++pending;
            else sDone[(isl >> 5) * kTile] |= 1u << (isl & 31);  // nothing to solve: minSeparation = 0 -> solved
            prev = e;
            prevJ = ej;
        }
    }
    while (!MbarTryWait(bodyBar, 0)) {}

    // Loop invariants pinned in registers and every ring / body-table access on 32-bit shared-window addresses, as in
    // k_solve_velocity: the chain is bound by the instructions the lone consumer warp has to issue per contact.
    const BodyTable T{Keep(SmemAddr(smem) + (uint32_t)lane * 4)};
    const uint32_t scratch = Keep(T.at(A.NB));
    const uint32_t stage0 = Keep(SmemAddr(stages) + (uint32_t)lane * 16);
    const uint32_t empty0 = Keep(SmemAddr(empty));
    const uint32_t full0 = Keep(SmemAddr(full));
    const uint32_t done0 = Keep(SmemAddr(sDone));
    const uint32_t ok0 = JOINTS ? Keep(SmemAddr(sDone + doneWords * kTile)) : 0u;
    const uint32_t mask = Keep((uint32_t)depth - 1u);
    const uint32_t depthShift = Keep((uint32_t)depthLog);
    const int nK = Keep(n), nMaxK = Keep(nMax);
    StsF(scratch, 0.0f); StsF(scratch + 128, 0.0f); StsF(scratch + 256, 0.0f);  // angle 0: the fast sinf / cosf path

    auto runPasses = [&](auto circlesTag, auto rotTag) {
        constexpr bool CIRCLES = decltype(circlesTag)::value;
        constexpr int ROT = decltype(rotTag)::value;
        int base = 0;  // ring index of this pass's record 0
        auto stageOf = [&](int k) { return stage0 + ((uint32_t)k & mask) * kPosStageBytes; };
        auto emptyOf = [&](int k) { return empty0 + (((uint32_t)k & mask) << 3); };
        auto fullOf = [&](int k) { return full0 + (((uint32_t)k & mask) << 3); };
        auto parityOf = [&](int k) { return ((uint32_t)k >> depthShift) & 1u; };
        uint32_t landedNext = 0;  // the test of record i + 2's barrier, issued a step ago
        auto doneBit = [&](int info) {  // converged bit of the island a record belongs to
            const uint32_t isl = (uint32_t)info >> 8;
            return ((LdsU(done0 + (isl >> 5) * (kTile * 4)) >> (isl & 31)) & 1u) != 0;
        };
        for (int pass = 0; pass < posIters; ++pass) {
            int i = 0;  // record being solved; records up to i + 1 are in registers
            float minSeparation = 0.0f;
            // record k of the pass (issued earlier) becomes usable
            auto finish = [&](PosRecord& r, int k) {
                FinishPosRecord(r, k < nK, T);
                __syncwarp();
                if (B2G_EARLY_RELEASE && k < nMaxK && lane == 0) MbarArriveA(emptyOf(base + k));
            };
            // One chain step: `cur` (record i of the pass) is solved on `bv` if `solve`; `fill` receives record i + 2 (past
            // the pass's last record: a stale stage, read and ignored) and the body values of `nxt` are read.
            auto step = [&](PosRecord& cur, const PosRecord& nxt, PosRecord& fill, BodyVals& bv, uint32_t& curA, uint32_t& curB,
                            bool& solve) {
                // record i + 2 (past the pass's last record there is nothing to wait for: the producer streams the next
                // pass only after this one's verdict)
                if (!landedNext && i + 2 < nMaxK) MbarSpinA(fullOf(base + i + 2), parityOf(base + i + 2));
                IssuePosRecord(fill, stageOf(base + i + 2));
                landedNext = MbarTestA(fullOf(base + i + 3), parityOf(base + i + 3));
                // a contact is solved if its island has not converged in an earlier pass
                const bool solveNext = (nxt.info & 8) != 0 && !doneBit(nxt.info);
                const uint32_t nxtA = solveNext ? nxt.pA : scratch, nxtB = solveNext ? nxt.pB : scratch;
                const BodyVals nextLoaded = LoadBodyVals(nxtA, nxtB);  // before this contact's stores: ForwardBodyVals
                {
                    const bool two = (cur.info & 16) != 0;
                    Vec2 cAp = V2(bv.ax, bv.ay), cBp = V2(bv.bx, bv.by);
                    float aA = bv.aw, aB = bv.bw;
                    bool slow = false;
                    const float s0 = SolvePositionPoint<CIRCLES, ROT>(cur, V2(cur.r1.x, cur.r1.y), cAp, aA, cBp, aB, slow);
                    Vec2 cAp2 = cAp, cBp2 = cBp;
                    float aA2 = aA, aB2 = aB;
                    const float s1 = SolvePositionPoint<CIRCLES, ROT>(cur, V2(cur.r1.z, cur.r1.w), cAp2, aA2, cBp2, aB2, slow);
                    BodyVals nb;
                    nb.ax = two ? cAp2.x : cAp.x; nb.ay = two ? cAp2.y : cAp.y; nb.aw = two ? aA2 : aA;
                    nb.bx = two ? cBp2.x : cBp.x; nb.by = two ? cBp2.y : cBp.y; nb.bw = two ? aB2 : aB;
                    float sep = two ? Min(s0, s1) : s0;  // Min(Min(m, s0), s1) == Min(m, Min(s0, s1)): no NaN reaches here unless the world is lost
                    if (slow && solve) {  // a body turned further than pi/4: the same again with the full-range sinf / cosf
                        const PosContactResult o = SolvePositionContactAnyAngle(cur.r0, cur.r1, cur.r2, cur.r3, cur.r4, cur.info, bv);
                        nb = o.bv;
                        sep = o.minSeparation;
                    }
                    minSeparation = solve ? Min(minSeparation, sep) : minSeparation;
                    // not solved: the scratch row keeps its zeros
                    bv.ax = solve ? nb.ax : bv.ax; bv.ay = solve ? nb.ay : bv.ay; bv.aw = solve ? nb.aw : bv.aw;
                    bv.bx = solve ? nb.bx : bv.bx; bv.by = solve ? nb.by : bv.by; bv.bw = solve ? nb.bw : bv.bw;
                }
                StoreBodyVals(curA, curB, bv);
                bv = ForwardBodyVals(nextLoaded, bv, nxtA, nxtB, curA, curB);
                if ((cur.info & 12) == 12) {  // last contact of its island: close the island
                    if (solve && minSeparation >= -3.0f * kLinearSlop) {
                        const uint32_t isl = (uint32_t)cur.info >> 8;
                        const uint32_t wd = (deferred ? ok0 : done0) + (isl >> 5) * (kTile * 4);
                        StsU(wd, LdsU(wd) | (1u << (isl & 31)));
                        if (!deferred) --pending;
                    }
                    minSeparation = 0.0f;
                }
                curA = nxtA;
                curB = nxtB;
                solve = solveNext;
                finish(fill, i + 2);
                // record i is solved and stored: its stage goes back (past the pass's last record there is none)
                if (!B2G_EARLY_RELEASE && i < nMaxK && lane == 0) MbarArriveA(emptyOf(base + i));
                ++i;
            };
            if (depth < 8 || B2G_POS_FORCE_SIMPLE) {
                // a handful of contacts per world: one record at a time, nothing overlapped
                for (; i < nMaxK; ++i) {
                    MbarSpinA(fullOf(base + i), parityOf(base + i));
                    PosRecord r;
                    IssuePosRecord(r, stageOf(base + i));
                    FinishPosRecord(r, i < nK, T);
                    const bool solve = (r.info & 8) != 0 && !doneBit(r.info);
                    if (solve) {
                        BodyVals bv = LoadBodyVals(r.pA, r.pB);
                        Vec2 cAp = V2(bv.ax, bv.ay), cBp = V2(bv.bx, bv.by);
                        float aA = bv.aw, aB = bv.bw;
                        bool slow = false;
                        minSeparation = Min(minSeparation, SolvePositionPoint<true, kRotAny>(r, V2(r.r1.x, r.r1.y), cAp, aA, cBp, aB, slow));
                        if (r.info & 16)
                            minSeparation = Min(minSeparation, SolvePositionPoint<true, kRotAny>(r, V2(r.r1.z, r.r1.w), cAp, aA, cBp, aB, slow));
                        bv.ax = cAp.x; bv.ay = cAp.y; bv.aw = aA;
                        bv.bx = cBp.x; bv.by = cBp.y; bv.bw = aB;
                        StoreBodyVals(r.pA, r.pB, bv);
                    }
                    __syncwarp();
                    if (lane == 0) MbarArriveA(emptyOf(base + i));
                    if ((r.info & 12) == 12) {
                        if (solve && minSeparation >= -3.0f * kLinearSlop) {
                            const uint32_t isl = (uint32_t)r.info >> 8;
                            const uint32_t wd = (deferred ? ok0 : done0) + (isl >> 5) * (kTile * 4);
                            StsU(wd, LdsU(wd) | (1u << (isl & 31)));
                            if (!deferred) --pending;
                        }
                        minSeparation = 0.0f;
                    }
                }
            } else {
            PosRecord r0, r1, r2;
            MbarSpinA(fullOf(base), parityOf(base));
            IssuePosRecord(r0, stageOf(base));
            finish(r0, 0);
            MbarSpinA(fullOf(base + 1), parityOf(base + 1));
            IssuePosRecord(r1, stageOf(base + 1));
            finish(r1, 1);
            landedNext = MbarTestA(fullOf(base + 2), parityOf(base + 2));
            bool solve = (r0.info & 8) != 0 && !doneBit(r0.info);
            uint32_t curA = solve ? r0.pA : scratch, curB = solve ? r0.pB : scratch;
            BodyVals bv = LoadBodyVals(curA, curB);
            // groups of three steps; the steps past the pass's last record have nothing to solve (and release no stage)
            while (i < nMaxK) {
                step(r0, r1, r2, bv, curA, curB, solve);
                step(r1, r2, r0, bv, curA, curB, solve);
                step(r2, r0, r1, bv, curA, curB, solve);
            }
            }
            base += nMaxK;
            if (anyJoints) {
                __syncwarp();
                if (deferred) {
                    const JointCols jc = MakeJointCols(A, tile, lane, 0, T.base);
                    int j0 = 0, c0 = 0;
                    for (int isl = 0; isl < islands; ++isl) {
                        const int j1 = iJointEnd[isl], c1 = iContactEnd[isl];
                        const uint32_t dw = done0 + ((uint32_t)isl >> 5) * (kTile * 4), ow = ok0 + ((uint32_t)isl >> 5) * (kTile * 4);
                        const uint32_t bit = 1u << (isl & 31);
                        if ((LdsU(dw) & bit) == 0) {
                            bool jointsOkay = true;
                            for (int j = j0; j < j1; ++j) {
                                const bool jointOkay = JointSolvePosition(&jc, iJoints[j]);
                                jointsOkay = jointsOkay && jointOkay;
                            }
                            const bool contactsOkay = c1 == c0 || (LdsU(ow) & bit) != 0;
                            if (contactsOkay && jointsOkay) {
                                StsU(dw, LdsU(dw) | bit);
                                --pending;
                            }
                            StsU(ow, LdsU(ow) & ~bit);
                        }
                        j0 = j1;
                        c0 = c1;
                    }
                }
                __syncwarp();
            }
            const bool more = pass + 1 < posIters && !__all_sync(0xffffffffu, pending == 0);
            if (lane == 0) *(volatile int*)&sContinue = more ? 1 : 0;
            __syncwarp();
            MbarArrive(passBar);
            if (!more) break;
        }
    };
    // Which arithmetic the tile's steps evaluate: has any body of any of its worlds turned further than pi/4 (anything that
    // rolls or tumbles)?  Then the wide-range sinf / cosf (and the e_circles arithmetic with it: one variant fewer).
    bool turned = false;
    {
        const int nb = A.bodyCount[ws];
        for (int b = 0; b < nb; ++b) turned |= !InSinCosFastRange(LdsF(T.at(b) + 256));
    }
#ifdef B2G_NO_WIDE_ROT  // A/B knob: turned bodies take the out-of-line full-range path, as before
    turned = false;
#endif
    const bool anyTurned = __any_sync(0xffffffffu, turned);
    if (anyTurned) runPasses(std::true_type(), std::integral_constant<int, kRotWide>());
    else if (anyCircles) runPasses(std::true_type(), std::integral_constant<int, kRotNarrow>());
    else runPasses(std::false_type(), std::integral_constant<int, kRotNarrow>());
    for (int wd = 0; wd < doneWords; ++wd) gDone[(size_t)wd * kTile] = (int)sDone[wd * kTile];
    FenceProxyAsync();
    __syncwarp();
    if (lane == 0) {
        TmaStore1D(gPos, smem, bodyBytes);
        TmaStoreCommitAndWait();
    }
}

// =================================================================================================
// k_solve_finalize: write-back + SynchronizeTransform, sleep
// =================================================================================================
__global__ void __launch_bounds__(kPrepWarps * 32) k_solve_finalize(Arena A, float h) {
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const bool live = tile * kTile + lane < A.numWorlds;
    const int nb = live ? A.bodyCount[ws] : 0;
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    Col<float4> bodyForce = col(A.bodyForce, tile, A.NB, lane);
    const float* sp = A.solvePos + (size_t)tile * A.NB * 3 * kTile + lane;
    const int nbMax = __reduce_max_sync(0xffffffffu, nb);
    for (int b = warp; b < nbMax; b += kPrepWarps) {
        if (b >= nb) continue;
        if ((bodyFlags[b] & kBodyIsland) == 0) continue;
        float4 sw = bodySweep[b];
        sw.x = sp[(size_t)(b * 3 + 0) * kTile];
        sw.y = sp[(size_t)(b * 3 + 1) * kTile];
        sw.z = sp[(size_t)(b * 3 + 2) * kTile];
        bodySweep[b] = sw;
        const float4 ms = bodyMass[b];
        Rot q = MakeRot(sw.z);
        Vec2 p = V2(sw.x, sw.y) - Mul(q, V2(ms.z, ms.w));
        bodyXf[b] = make_float4(p.x, p.y, q.s, q.c);
    }
    const bool allowSleep = live && (A.worldFlags[ws] & B2G_WORLD_ALLOW_SLEEP) != 0;
    if (!__any_sync(0xffffffffu, allowSleep)) return;
    __syncthreads();
    // ---- sleep (b2Island.cpp:345-382) ----
    // Pass 1, all warps, entries of the island body lists strided over warps: every island body's sleep time is advanced
    // or reset, and what it contributes to its island's minSleepTime goes to a per-entry scratch column (the velocity
    // records are dead by now).  Pass 2, warp 0: per island the minimum over its entries — one coalesced stream of
    // independent loads instead of a chain of dependent ones per body — and, rarely, the island goes to sleep.
    const int islands = allowSleep ? A.islandCount[ws] : 0;
    Col<int> iBodies = col(A.islandBodies, tile, A.NB + A.NC, lane);
    Col<int> iBodyEnd = col(A.islandBodyEnd, tile, A.NB, lane);
    float* contribution = reinterpret_cast<float*>(A.velRec + (size_t)tile * A.NC * kVelRec * kTile) + lane;  // [entry][lane]
    const int entries = islands > 0 ? iBodyEnd[islands - 1] : 0;
    const int entriesMax = __reduce_max_sync(0xffffffffu, entries);
    const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
    const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
    for (int i = warp; i < entriesMax; i += kPrepWarps) {
        if (i >= entries) continue;
        const int b = iBodies[i];
        const int f = bodyFlags[b];
        float c = -1.0f;  // a static body: no part in the minimum; pass 2 re-wakes it in island order
        if ((f & kBodyTypeMask) != B2G_STATIC_BODY) {
            const float4 vel = bodyVel[b];
            float4 sw = bodySweep[b];
            if ((f & kBodyAutoSleep) == 0 || vel.z * vel.z > angTolSqr || Dot(V2(vel.x, vel.y), V2(vel.x, vel.y)) > linTolSqr) {
                sw.w = 0.0f;
            } else {
                sw.w += h;
            }
            bodySweep[b] = sw;
            c = sw.w;
        }
        contribution[(size_t)i * kTile] = c;
    }
    __syncthreads();
    if (warp != 0 || !allowSleep) return;
    const int doneWords = (A.NB + 31) / 32;
    const int* gDone = A.islandSolved + (size_t)tile * doneWords * kTile + lane;
    int b0 = 0;
    for (int isl = 0; isl < islands; ++isl) {
        const int b1 = iBodyEnd[isl];
        float minSleepTime = kMaxFloat;
        auto take = [&](int i, float c) {
            if (c < 0.0f) {
                // b2World::Solve builds and solves islands alternately: this island's DFS re-woke a static body
                // that an earlier island of the same step had put to sleep (b2World.cpp:441)
                const int b = iBodies[i];
                const int f = bodyFlags[b];
                if ((f & kBodyAwake) == 0) bodyFlags[b] = f | kBodyAwake;
            } else {
                minSleepTime = Min(minSleepTime, c);
            }
        };
        int i = b0;
        for (; i + 4 <= b1; i += 4) {
            const float c0 = contribution[(size_t)i * kTile], c1 = contribution[(size_t)(i + 1) * kTile];
            const float c2 = contribution[(size_t)(i + 2) * kTile], c3 = contribution[(size_t)(i + 3) * kTile];
            take(i, c0); take(i + 1, c1); take(i + 2, c2); take(i + 3, c3);
        }
        for (; i < b1; ++i) take(i, contribution[(size_t)i * kTile]);
        const bool positionSolved = (gDone[(size_t)(isl >> 5) * kTile] >> (isl & 31) & 1) != 0;
        if (minSleepTime >= kTimeToSleep && positionSolved) {
            A.topoDirty[ws] = 1;
            for (int k = b0; k < b1; ++k) {  // b2Body::SetAwake(false), b2Body.h:644-652
                int b = iBodies[k];
                bodyFlags[b] &= ~kBodyAwake;
                float4 sw = bodySweep[b];
                sw.w = 0.0f;
                bodySweep[b] = sw;
                bodyVel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        b0 = b1;
    }
}

}  // namespace b2g
