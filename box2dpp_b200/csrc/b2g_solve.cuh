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

#include "b2g_kernels.cuh"

namespace b2g {

constexpr int kVelRec = 7;      // float4 per velocity record
constexpr int kPosRec = 5;      // float4 per position record
#ifndef B2G_REC_STAGES
#define B2G_REC_STAGES 6
#endif
constexpr int kRecStages = B2G_REC_STAGES;   // TMA ring depth of the position kernel (records in flight per tile)
#ifndef B2G_VEL_STAGES
#define B2G_VEL_STAGES 7
#endif
constexpr int kVelStages = B2G_VEL_STAGES;   // TMA ring depth of the velocity kernel
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

__device__ __forceinline__ int PackBodies(int bodyA, int bodyB, int pointCount) { return bodyA | (bodyB << 15) | (pointCount << 30); }

// Shared-memory budget of the sequential kernels for a world capacity of NB bodies.
__host__ __device__ inline size_t SolveSmemBytes(int NB) {
    size_t body = ((size_t)NB * 3 * kTile * 4 + 127) & ~(size_t)127;
    size_t done = (size_t)((NB + 31) / 32) * kTile * 4;
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
    for (int i = warp; i < nMax; i += kPrepWarps) {
        if (i >= n) continue;
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
        pr[4 * kTile] = make_float4(radiusA, radiusB, __int_as_float(PackBodies(bodyA, bodyB, pointCount)), __int_as_float(mtype));
    }
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
    // suspend-time hint: the producer / watcher warps sleep in hardware until the phase completes instead of
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
__device__ __forceinline__ uint32_t LdsAcquire(uint32_t a) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void StsRelease(uint32_t a, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
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

struct VelRecord {
    float4 q0, q1, q2, q3, q4, q5, q6, imp;
};
__device__ __forceinline__ VelRecord LoadVelRecord(uint32_t st) {  // st: shared address of this lane's float4 0
    VelRecord r;
    r.q0 = LdsF4(st + 0 * 512); r.q1 = LdsF4(st + 1 * 512); r.q2 = LdsF4(st + 2 * 512); r.q3 = LdsF4(st + 3 * 512);
    r.q4 = LdsF4(st + 4 * 512); r.q5 = LdsF4(st + 5 * 512); r.q6 = LdsF4(st + 6 * 512); r.imp = LdsF4(st + 7 * 512);
    return r;
}

// WarmStart of one contact (b2ContactSolver.cpp:240-278)
__device__ __forceinline__ void WarmStartContact(const VelRecord& r, BodyTable T) {
    const int packed = __float_as_int(r.q0.w);
    const int bodyA = packed & 0x7fff, bodyB = (packed >> 15) & 0x7fff;
    const int pointCount = (packed >> 30) & 3;
    const float mA = r.q1.x, iA = r.q1.y, mB = r.q1.z, iB = r.q1.w;
    const uint32_t pA = T.at(bodyA), pB = T.at(bodyB);
    Vec2 vA = V2(LdsF(pA), LdsF(pA + 128)), vB = V2(LdsF(pB), LdsF(pB + 128));
    float wA = LdsF(pA + 256), wB = LdsF(pB + 256);
    const Vec2 normal = V2(r.q0.x, r.q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    {
        Vec2 P = r.imp.x * normal + r.imp.y * tangent;
        wA -= iA * Cross(V2(r.q2.x, r.q2.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(r.q2.z, r.q2.w), P);
        vB = vB + mB * P;
    }
    if (pointCount > 1) {
        Vec2 P = r.imp.z * normal + r.imp.w * tangent;
        wA -= iA * Cross(V2(r.q4.x, r.q4.y), P);
        vA = vA - mA * P;
        wB += iB * Cross(V2(r.q4.z, r.q4.w), P);
        vB = vB + mB * P;
    }
    StsF(pA, vA.x); StsF(pA + 128, vA.y); StsF(pA + 256, wA);
    StsF(pB, vB.x); StsF(pB + 128, vB.y); StsF(pB + 256, wB);
}

// SolveVelocityConstraints for one contact (b2ContactSolver.cpp:280-597); returns the updated impulses
__device__ __forceinline__ float4 SolveVelocityContact(const VelRecord& r, BodyTable T) {
    const float4 q0 = r.q0, q1 = r.q1, q2 = r.q2, q3 = r.q3;
    float4 imp = r.imp;
    const int packed = __float_as_int(q0.w);
    const int bodyA = packed & 0x7fff, bodyB = (packed >> 15) & 0x7fff;
    const int pointCount = (packed >> 30) & 3;
    const float mA = q1.x, iA = q1.y, mB = q1.z, iB = q1.w;
    const uint32_t pA = T.at(bodyA), pB = T.at(bodyB);
    Vec2 vA = V2(LdsF(pA), LdsF(pA + 128)), vB = V2(LdsF(pB), LdsF(pB + 128));
    float wA = LdsF(pA + 256), wB = LdsF(pB + 256);
    const Vec2 normal = V2(q0.x, q0.y);
    const Vec2 tangent = Cross(normal, 1.0f);
    const Vec2 rA0 = V2(q2.x, q2.y), rB0 = V2(q2.z, q2.w);
    const float friction = q0.z;
    {  // tangent constraint of point 0 (tangentSpeed is 0: b2Contact::SetTangentSpeed is not exposed)
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
        const float4 q4 = r.q4, q5 = r.q5, q6 = r.q6;
        const Vec2 rA1 = V2(q4.x, q4.y), rB1 = V2(q4.z, q4.w);
        {  // tangent constraint of point 1
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
        // block solver (b2ContactSolver.cpp:388-589).  The reference tries four cases in order and takes the
        // first feasible one; here all four are evaluated branch-free and the same one is selected.
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
        if (solved) {
            vA = nvA; wA = nwA; vB = nvB; wB = nwB;
            imp.x = x.x;
            imp.z = x.y;
        }
    }
    StsF(pA, vA.x); StsF(pA + 128, vA.y); StsF(pA + 256, wA);
    StsF(pB, vB.x); StsF(pB + 128, vB.y); StsF(pB + 256, wB);
    return imp;
}

__global__ void __launch_bounds__(96) k_solve_velocity(Arena A, int velIters) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t sReady;  // records whose data has landed in the ring (published by the watcher warp)
    const int tile = A.tile0 + blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t ws = (size_t)tile * kTile + lane;
    const int n = A.solveCount[ws];
    const int nMax = __reduce_max_sync(0xffffffffu, n);
    if (nMax == 0) return;  // no touching contact in the whole tile: velocities stay as integrated
    const bool warm = (A.worldFlags[ws] & B2G_WORLD_WARM_STARTING) != 0;
    const int anyWarm = __any_sync(0xffffffffu, warm) ? 1 : 0;
    const int total = nMax * (velIters + anyWarm);
    // Ring depth.  The accumulated impulses of a contact travel with its record: the consumer leaves the updated
    // values in the ring stage, and the producer lane that fetched them (cp.async) writes them back to global
    // memory when the stage is recycled and fetches them again for the next pass — store and re-fetch of one
    // contact are issued by the same thread, in that order, as long as the ring holds fewer than nMax records.
    const int depth = nMax - 1 < 1 ? 1 : (nMax - 1 < kVelStages ? nMax - 1 : kVelStages);

    const uint32_t bodyBytes = (uint32_t)A.NB * 3 * kTile * 4;
    unsigned char* stages = smem + ((bodyBytes + 127) & ~127u);
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + kVelStages * kVelStageBytes);
    uint64_t* empty = full + kVelStages;
    uint64_t* bodyBar = empty + kVelStages;
    float* gVel = A.solveVel + (size_t)tile * A.NB * 3 * kTile;
    float4* gImp = A.impRec + (size_t)tile * A.NC * kTile + lane;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kVelStages; ++s) {
            MbarInit(&full[s], 33);   // expect_tx arrive of the TMA issuer + 32 cp.async completions
            MbarInit(&empty[s], 32);  // every consumer lane
        }
        MbarInit(bodyBar, 1);
        sReady = 0;
        FenceBarrierInit();
    }
    __syncthreads();

    if (warp == 2) {
        // ================= watcher: waits for every record in order and publishes the count =================
        // (mbarrier.try_wait returns a predicate that must be consumed at once, ~50-90 cycles; the consumer
        // polls this word with an ordinary load instead, whose latency it overlaps with the solve)
        if (lane != 0) return;
        const uint32_t ready = SmemAddr(&sReady);
        int s = 0;
        uint32_t phase = 0;
        for (int r = 0; r < total; ++r) {
            while (!MbarTryWait(&full[s], phase)) {}
            StsRelease(ready, (uint32_t)r + 1u);
            if (++s == depth) { s = 0; phase ^= 1u; }
        }
        return;
    }
    if (warp == 1) {
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
        for (int seq = 0; seq < total + depth; ++seq) {
            float4* st = reinterpret_cast<float4*>(stages + (size_t)s * kVelStageBytes) + lane;
            if (seq >= depth) {  // stage s holds record seq - depth: wait for the consumer, write its impulses back
                while (!MbarTryWait(&empty[s], phase)) {}
                if (wi < n && !(anyWarm && wpass == 0)) gImp[(size_t)wi * kTile] = st[7 * kTile];
                if (fence) __threadfence();
                if (++wi == nMax) { wi = 0; ++wpass; }
            }
            if (seq < total) {
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
        return;
    }

    // ================= consumer =================
    while (!MbarTryWait(bodyBar, 0)) {}
    // Loop invariants are pinned in registers (Keep): left alone, the compiler re-derives them from %tid / %ctaid /
    // the shared-window base inside the loop, which costs more than the solve's own address arithmetic.
    const BodyTable T{Keep(SmemAddr(smem) + (uint32_t)lane * 4)};
    const uint32_t stage0 = Keep(SmemAddr(stages) + (uint32_t)lane * 16);
    const uint32_t empty0 = Keep(SmemAddr(empty));
    const uint32_t ready = Keep(SmemAddr(&sReady));
    const uint32_t stageEnd = Keep(stage0 + (uint32_t)depth * kVelStageBytes);
    const int warmRecords = Keep(anyWarm ? nMax : 0);
    const int nK = Keep(n), nMaxK = Keep(nMax), totalK = Keep(total);
    if (depth >= 3) {
        // Two records of look-ahead in registers: the record solved in an iteration was loaded one iteration
        // earlier, so neither the ring's barrier nor the LDS latency sits on the solver's dependency chain.
        uint32_t ldStage = stage0, stStage = stage0, stEmpty = empty0;
        int i = 0, seq = 0;
        auto load = [&](VelRecord& r) {
            r = LoadVelRecord(ldStage);
            ldStage += kVelStageBytes;
            if (ldStage == stageEnd) ldStage = stage0;
        };
        // one solver step on the record held in `r`, then refill `r` with the record two ahead
        auto step = [&](VelRecord& r) {
            uint32_t landed = LdsAcquire(ready);  // polled now, needed after the solve
            if (i < nK) {
                if (seq < warmRecords) {
                    if (warm) WarmStartContact(r, T);
                } else {
                    StsF4(stStage + 7 * 512, SolveVelocityContact(r, T));
                }
            }
            MbarArriveA(stEmpty);  // hands the stage (and the updated impulses in it) back to the producer
            stStage += kVelStageBytes;
            stEmpty += 8;
            if (stStage == stageEnd) { stStage = stage0; stEmpty = empty0; }
            if (seq + 2 < totalK) {
                while (landed < (uint32_t)seq + 3u) landed = LdsAcquire(ready);
                load(r);
            }
            if (++i == nMaxK) i = 0;
            ++seq;
        };
        VelRecord r0, r1;
        while (LdsAcquire(ready) < 2u) {}
        load(r0);
        load(r1);
        while (seq < totalK) {
            step(r0);
            if (seq < totalK) step(r1);
        }
    } else {
        int i = 0, cs = 0;
        for (int seq = 0; seq < total; ++seq) {
            const uint32_t st = stage0 + cs * kVelStageBytes;
            while (LdsAcquire(ready) < (uint32_t)seq + 1u) {}
            if (i < n) {
                const VelRecord r = LoadVelRecord(st);
                if (seq < warmRecords) {
                    if (warm) WarmStartContact(r, T);
                } else {
                    StsF4(st + 7 * 512, SolveVelocityContact(r, T));
                }
            }
            MbarArriveA(empty0 + cs * 8);
            if (++i == nMax) i = 0;
            if (++cs == depth) cs = 0;
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
// consumer's verdict on whether another pass is needed.
// =================================================================================================
constexpr int kPosStageBytes = kPosRec * 16 * kTile;  // 2560

__global__ void __launch_bounds__(64) k_solve_position(Arena A, int posIters) {
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
    if (nMax == 0 || posIters <= 0) {
        // islands without contacts converge in the first pass (minSeparation == 0); no pass, no convergence
        if (warp == 0)
            for (int wd = 0; wd < doneWords; ++wd) gDone[(size_t)wd * kTile] = posIters > 0 ? -1 : 0;
        return;
    }

    const uint32_t bodyBytes = (uint32_t)A.NB * 3 * kTile * 4;
    unsigned char* stages = smem + ((bodyBytes + 127) & ~127u);
    uint32_t* sDone = reinterpret_cast<uint32_t*>(stages + kRecStages * kPosStageBytes) + lane;  // word wd at sDone[wd * 32]
    uint64_t* full = reinterpret_cast<uint64_t*>(stages + kRecStages * kPosStageBytes + (size_t)doneWords * kTile * 4);
    uint64_t* empty = full + kRecStages;
    uint64_t* bodyBar = empty + kRecStages;
    uint64_t* passBar = bodyBar + 1;
    float* gPos = A.solvePos + (size_t)tile * A.NB * 3 * kTile;
    const int depth = nMax < kRecStages ? nMax : kRecStages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRecStages; ++s) {
            MbarInit(&full[s], 1);    // the TMA issuer's expect_tx arrive
            MbarInit(&empty[s], 32);  // every consumer lane
        }
        MbarInit(bodyBar, 1);
        MbarInit(passBar, 32);
        FenceBarrierInit();
    }
    __syncthreads();

    if (warp == 1) {
        // ================= producer (one lane) =================
        if (lane != 0) return;
        MbarExpectTx(bodyBar, bodyBytes);
        TmaLoad1D(smem, gPos, bodyBytes, bodyBar);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(A.posRec + (size_t)tile * A.NC * kPosRec * kTile);
        int s = 0, use = 0;
        for (int pass = 0; pass < posIters; ++pass) {
            for (int i = 0; i < nMax; ++i) {
                if (use > 0) while (!MbarTryWait(&empty[s], (uint32_t)(use - 1) & 1u)) {}
                MbarExpectTx(&full[s], kPosStageBytes);
                TmaLoad1D(stages + (size_t)s * kPosStageBytes, src + (size_t)i * kPosStageBytes, kPosStageBytes, &full[s]);
                if (++s == depth) { s = 0; ++use; }
            }
            while (!MbarTryWait(passBar, (uint32_t)pass & 1u)) {}
            if (*(volatile int*)&sContinue == 0) break;
        }
        return;
    }

    // ================= consumer =================
    for (int wd = 0; wd < doneWords; ++wd) sDone[wd * kTile] = 0u;
    // islands that own at least one contact and have not converged yet
    int pending = 0;
    {
        int prev = 0;
        for (int isl = 0; isl < islands; ++isl) {
            int e = iContactEnd[isl];
            if (e > prev) ++pending;
            else sDone[(isl >> 5) * kTile] |= 1u << (isl & 31);  // no contacts: minSeparation = 0 -> solved
            prev = e;
        }
    }
    while (!MbarTryWait(bodyBar, 0)) {}

    // Loop invariants pinned in registers and every ring / body-table access on 32-bit shared-window addresses, as in
    // k_solve_velocity: the chain is bound by the instructions the lone consumer warp has to issue per contact.
    const BodyTable T{Keep(SmemAddr(smem) + (uint32_t)lane * 4)};
    const uint32_t stage0 = Keep(SmemAddr(stages) + (uint32_t)lane * 16);
    const uint32_t stageEnd = Keep(stage0 + (uint32_t)depth * kPosStageBytes);
    const uint32_t full0 = Keep(SmemAddr(full)), empty0 = Keep(SmemAddr(empty));
    const int nK = Keep(n), nMaxK = Keep(nMax);
    uint32_t st = stage0, stFull = full0, stEmpty = empty0, parity = 0;
    for (int pass = 0; pass < posIters; ++pass) {
        int isl = 0;
        int islEnd = islands > 0 ? iContactEnd[0] : 0;
        float minSeparation = 0.0f;
        bool islDone = islands > 0 ? ((sDone[0] & 1u) != 0) : true;
        for (int i = 0; i < nMaxK; ++i) {
            while (!MbarTryWaitA(stFull, parity)) {}
            if (i < nK) {
                while (i >= islEnd) {  // entering the next island: close the previous one
                    if (!islDone && minSeparation >= -3.0f * kLinearSlop) {
                        sDone[(isl >> 5) * kTile] |= 1u << (isl & 31);
                        --pending;
                    }
                    ++isl;
                    islEnd = iContactEnd[isl];
                    minSeparation = 0.0f;
                    islDone = (sDone[(isl >> 5) * kTile] >> (isl & 31) & 1u) != 0;
                }
                if (!islDone) {
                    const float4 r0 = LdsF4(st), r1 = LdsF4(st + 512), r2 = LdsF4(st + 1024), r3 = LdsF4(st + 1536),
                                 r4 = LdsF4(st + 2048);
                    const int packed = __float_as_int(r4.z);
                    const int bodyA = packed & 0x7fff, bodyB = (packed >> 15) & 0x7fff;
                    const int pointCount = (packed >> 30) & 3;
                    const int mtype = __float_as_int(r4.w);
                    const float radiusA = r4.x, radiusB = r4.y;
                    const float mA = r2.x, iA = r2.y, mB = r2.z, iB = r2.w;
                    const Vec2 localCenterA = V2(r3.x, r3.y), localCenterB = V2(r3.z, r3.w);
                    const Vec2 localNormal = V2(r0.x, r0.y), localPoint = V2(r0.z, r0.w);
                    const uint32_t pA = T.at(bodyA), pB = T.at(bodyB);
                    Vec2 cAp = V2(LdsF(pA), LdsF(pA + 128)), cBp = V2(LdsF(pB), LdsF(pB + 128));
                    float aA = LdsF(pA + 256), aB = LdsF(pB + 256);
                    // one manifold point (b2ContactSolver.cpp:687-734)
                    auto solvePoint = [&](const Vec2 lp) {
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
                        } else {
                            // e_faceA / e_faceB (b2ContactSolver.cpp:642-664) are the same arithmetic with the bodies'
                            // roles swapped: selected, not branched on — lane = world, so one warp holds both kinds
                            const bool faceA = mtype == kManifoldFaceA;
                            Xf xfR, xfI;  // the reference face's body, the incident one
                            xfR.q.s = faceA ? xfA.q.s : xfB.q.s; xfR.q.c = faceA ? xfA.q.c : xfB.q.c;
                            xfR.p.x = faceA ? xfA.p.x : xfB.p.x; xfR.p.y = faceA ? xfA.p.y : xfB.p.y;
                            xfI.q.s = faceA ? xfB.q.s : xfA.q.s; xfI.q.c = faceA ? xfB.q.c : xfA.q.c;
                            xfI.p.x = faceA ? xfB.p.x : xfA.p.x; xfI.p.y = faceA ? xfB.p.y : xfA.p.y;
                            normal = Mul(xfR.q, localNormal);
                            Vec2 planePoint = Mul(xfR, localPoint);
                            Vec2 clipPoint = Mul(xfI, lp);
                            separation = Dot(clipPoint - planePoint, normal) - radiusA - radiusB;
                            point = clipPoint;
                            normal.x = faceA ? normal.x : -normal.x;
                            normal.y = faceA ? normal.y : -normal.y;
                        }
                        Vec2 rA = point - cAp;
                        Vec2 rB = point - cBp;
                        minSeparation = Min(minSeparation, separation);
                        float C = Clamp(kBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
                        float rnA = Cross(rA, normal);
                        float rnB = Cross(rB, normal);
                        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
                        float impulse = K > 0.0f ? Div(-C, K) : 0.0f;
                        Vec2 P = impulse * normal;
                        cAp = cAp - mA * P;
                        aA -= iA * Cross(rA, P);
                        cBp = cBp + mB * P;
                        aB += iB * Cross(rB, P);
                    };
                    if (pointCount > 0) solvePoint(V2(r1.x, r1.y));
                    if (pointCount > 1) solvePoint(V2(r1.z, r1.w));
                    StsF(pA, cAp.x); StsF(pA + 128, cAp.y); StsF(pA + 256, aA);
                    StsF(pB, cBp.x); StsF(pB + 128, cBp.y); StsF(pB + 256, aB);
                }
            }
            MbarArriveA(stEmpty);
            st += kPosStageBytes;
            stFull += 8;
            stEmpty += 8;
            if (st == stageEnd) { st = stage0; stFull = full0; stEmpty = empty0; parity ^= 1u; }
        }
        if (n > 0 && !islDone && minSeparation >= -3.0f * kLinearSlop) {  // close the last island of this pass
            sDone[(isl >> 5) * kTile] |= 1u << (isl & 31);
            --pending;
        }
        const bool more = pass + 1 < posIters && !__all_sync(0xffffffffu, pending == 0);
        if (lane == 0) *(volatile int*)&sContinue = more ? 1 : 0;
        __syncwarp();
        MbarArrive(passBar);
        if (!more) break;
    }
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
    if (warp != 0 || !allowSleep) return;
    // ---- sleep (b2Island.cpp:345-382): islands one after another, bodies in island order ----
    const int islands = A.islandCount[ws];
    Col<int> iBodies = col(A.islandBodies, tile, A.NB + A.NC, lane);
    Col<int> iBodyEnd = col(A.islandBodyEnd, tile, A.NB, lane);
    const int doneWords = (A.NB + 31) / 32;
    const int* gDone = A.islandSolved + (size_t)tile * doneWords * kTile + lane;
    const float linTolSqr = kLinearSleepTolerance * kLinearSleepTolerance;
    const float angTolSqr = kAngularSleepTolerance * kAngularSleepTolerance;
    int b0 = 0;
    for (int isl = 0; isl < islands; ++isl) {
        const int b1 = iBodyEnd[isl];
        float minSleepTime = kMaxFloat;
        for (int i = b0; i < b1; ++i) {
            int b = iBodies[i];
            int f = bodyFlags[b];
            if ((f & kBodyTypeMask) == B2G_STATIC_BODY) {
                // b2World::Solve builds and solves islands alternately: this island's DFS re-woke a static body
                // that an earlier island of the same step had put to sleep (b2World.cpp:441)
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
        const bool positionSolved = (gDone[(size_t)(isl >> 5) * kTile] >> (isl & 31) & 1) != 0;
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
        b0 = b1;
    }
}

}  // namespace b2g
