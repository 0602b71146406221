// b2g_step.cu — device layer of libb2gpu.so: arena allocation, field transfer, kernel launches and the
// b2gpu_batch_* C ABI declared in include/b2gpu.h.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// --fmad=false -prec-div=true -prec-sqrt=true -ftz=false (see __graft_entry__.py).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b2gpu.h"
#include "b2g_kernels.cuh"
#include "b2g_solve.cuh"
#include "b2g_toi.cuh"
#include "b2g_query.cuh"

using namespace b2g;

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) return fail(B2GPU_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        return e == cudaSuccess ? B2GPU_OK : fail(B2GPU_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
};

struct Field {
    void** dev;      // address of the arena pointer
    int capacity;    // items per world
    int components;  // 32-bit words per item
};

// world-major host layout <-> tile-interleaved arena; one thread per (world, item, component)
__global__ void k_interleave(uint32_t* __restrict__ arena, const uint32_t* __restrict__ flat, int cap, int comps,
                             int firstWorld, int worldCount, int toArena) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)worldCount * cap * comps;
    if (idx >= total) return;
    int c = (int)(idx % comps);
    size_t r = idx / comps;
    int item = (int)(r % cap);
    int wl = (int)(r / cap);
    int w = firstWorld + wl;
    size_t a = (((size_t)(w >> 5) * cap + item) * kTile + (w & 31)) * comps + c;
    if (toArena) arena[a] = flat[idx];
    else const_cast<uint32_t*>(flat)[idx] = arena[a];
}

// b2Body::ApplyForceToCenter / ApplyTorque for every body (Box2D/Dynamics/b2Body.h:738-790):
// wake on a non-zero force, then accumulate if awake.
// Covers tiles [A.tile0, A.tile0 + tileCount): the whole batch, or one world group's share (b2gpu_batch_step_io).
// `sel` (device, nSel body indices, the same in every world) picks the bodies that take part; null = every body slot.
__global__ void k_apply_forces(Arena A, const float* __restrict__ forces, int tileCount, const int* __restrict__ sel, int nSel) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // lane-fastest: (tile, selected body, lane)
    size_t total = (size_t)tileCount * nSel * kTile;
    if (idx >= total) return;
    int lane = (int)(idx & 31);
    size_t r = idx >> 5;
    int k = (int)(r % nSel);
    int tile = A.tile0 + (int)(r / nSel);
    int body = sel ? sel[k] : k;
    int w = tile * kTile + lane;
    if (w >= A.numWorlds || body >= A.bodyCount[(size_t)tile * kTile + lane]) return;
    const float* f = forces + ((size_t)w * nSel + k) * 3;
    float fx = f[0], fy = f[1], tq = f[2];
    if (fx == 0.0f && fy == 0.0f && tq == 0.0f) return;
    idx = ((size_t)tile * A.NB + body) * kTile + lane;
    int flags = A.bodyFlags[idx];
    if ((flags & kBodyTypeMask) != B2G_DYNAMIC_BODY || (flags & kBodyDestroyed)) return;
    if ((flags & kBodyAwake) == 0) {
        A.bodyFlags[idx] = flags | kBodyAwake;
        A.bodySweep[idx].w = 0.0f;
        A.topoDirty[(size_t)tile * kTile + lane] = 1;
    }
    float4 cur = A.bodyForce[idx];
    A.bodyForce[idx] = make_float4(cur.x + fx, cur.y + fy, cur.z + tq, 0.0f);
}

// b2Body::GetTransform (format B2GPU_POSE_XF: p.x p.y q.s q.c) or GetPosition + GetAngle (B2GPU_POSE_XYA: p.x p.y angle)
// of the selected bodies, world-major output
__global__ void k_read_transforms(Arena A, float* __restrict__ out, int tileCount, const int* __restrict__ sel, int nSel, int format) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)tileCount * nSel * kTile;
    if (idx >= total) return;
    int lane = (int)(idx & 31);
    size_t r = idx >> 5;
    int k = (int)(r % nSel);
    int tile = A.tile0 + (int)(r / nSel);
    int body = sel ? sel[k] : k;
    int w = tile * kTile + lane;
    if (w >= A.numWorlds) return;
    idx = ((size_t)tile * A.NB + body) * kTile + lane;
    const float4 xf = A.bodyXf[idx];
    if (format == B2GPU_POSE_XYA) {
        float* o = out + ((size_t)w * nSel + k) * 3;
        o[0] = xf.x; o[1] = xf.y; o[2] = A.bodySweep[idx].z;
    } else {
        reinterpret_cast<float4*>(out)[(size_t)w * nSel + k] = xf;
    }
}

// b2Body state for every body, world-major output: p.x p.y angle awake | v.x v.y omega sleepTime
__global__ void k_read_body_state(Arena A, float4* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)A.numTiles * A.NB * kTile;
    if (idx >= total) return;
    int lane = (int)(idx & 31);
    size_t r = idx >> 5;
    int body = (int)(r % A.NB);
    int tile = (int)(r / A.NB);
    int w = tile * kTile + lane;
    if (w >= A.numWorlds) return;
    const float4 xf = A.bodyXf[idx], sw = A.bodySweep[idx], v = A.bodyVel[idx];
    const int f = A.bodyFlags[idx];
    float4* o = out + ((size_t)w * A.NB + body) * 2;
    o[0] = make_float4(xf.x, xf.y, sw.z, (f & kBodyAwake) ? 1.0f : 0.0f);
    o[1] = make_float4(v.x, v.y, v.z, sw.w);
}

// ---- building blocks of the between-step edits (k_body_commands, k_attach_fixtures, k_create_joints) ----

// b2Contact::FlagForFiltering on every contact between two bodies (b2World.cpp:255-263, 342-358)
__device__ __noinline__ void FlagContactsBetween(const Arena& A, int tile, int lane, int bodyA, int bodyB) {
    Col<int4> cA = col(A.contactA, tile, A.NC, lane);
    const int nc = A.contactCount[(size_t)tile * kTile + lane];
    const int ab = bodyA | (bodyB << 16), ba = bodyB | (bodyA << 16);
    for (int k = 0; k < nc; ++k) {
        int4 a = cA[k];
        if (a.y != ab && a.y != ba) continue;
        a.z |= kContactFilter;
        cA[k] = a;
    }
}

// b2ContactManager::Destroy (EndContact if touching, b2ContactManager.cpp:45-105) + b2Contact::Destroy (wake both bodies if
// the manifold had points, b2Contact.cpp:103-122) for the contacts of body b — only those of fixture `fixture` when it
// is >= 0 — in the body's contact-list order (newest first); then a stable erase keeps every list's order.
__device__ __noinline__ void DestroyContactsOf(const Arena& A, int tile, int lane, int b, int fixture) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int4> cA = col(A.contactA, tile, A.NC, lane), cK = col(A.contactK, tile, A.NC, lane);
    Col<int2> cS = col(A.contactS, tile, A.NC, lane);
    Col<float4> cP0 = col(A.contactP0, tile, A.NC, lane), cP1 = col(A.contactP1, tile, A.NC, lane);
    Col<float4> cN = col(A.contactN, tile, A.NC, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    EventSink ev{col(A.eventA, tile, A.NE, lane), col(A.eventB, tile, A.NE, lane), &A.eventCount[ws], &A.status[ws], A.NE, A.eventMask};
    const int nc = A.contactCount[ws];
    bool any = false;
    for (int k = nc - 1; k >= 0; --k) {
        int4 a = cA[k];
        const int kA = a.y & 0xffff, kB = (a.y >> 16) & 0xffff;
        if (kA != b && kB != b) continue;
        const int pa = proxyInfo[a.x & 0xffff], pb = proxyInfo[(a.x >> 16) & 0xffff];
        if (fixture >= 0 && (pa & 0xffff) != fixture && (pb & 0xffff) != fixture) continue;
        any = true;
        const int pointCount = (a.w >> 8) & 0xff;
        if (a.z & kContactTouching)
            ev.push(kEventEnd, pa & 0xffff, (pa >> 16) & 0xffff, pb & 0xffff, (pb >> 16) & 0xffff, pointCount, make_float4(0, 0, 0, 0));
        if (pointCount > 0 && (a.z & kContactSensor) == 0) {
            WakeBody(bodyFlags, bodySweep, kA);
            WakeBody(bodyFlags, bodySweep, kB);
        }
        a.z |= kContactDead;
        cA[k] = a;
    }
    if (!any) return;
    A.topoDirty[ws] = 1;
    int wpos = 0;
    for (int k = 0; k < nc; ++k) {
        const int4 a = cA[k];
        if (a.z & kContactDead) continue;
        if (wpos != k) { cA[wpos] = a; cK[wpos] = cK[k]; cS[wpos] = cS[k]; cP0[wpos] = cP0[k]; cP1[wpos] = cP1[k]; cN[wpos] = cN[k]; }
        ++wpos;
    }
    A.contactCount[ws] = wpos;
}

// b2Fixture::DestroyProxies (b2Fixture.cpp:141-152): an empty fat AABB never overlaps anything, the proxy leaves the move
// buffer (b2BroadPhase::UnBufferMove) and its slot goes onto the LIFO free stack (b2DynamicTree::FreeNode), children ascending.
__device__ __noinline__ void DestroyFixtureProxies(const Arena& A, int tile, int lane, int fx) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    Col<int> proxyFree = col(A.proxyFree, tile, A.NP, lane);
    const int np = A.proxyCount[ws];
    int nfree = A.proxyFreeCount[ws];
    // the proxies of one fixture were allocated in ascending child order, but (after slot reuse) not in ascending slots
    int children = 0;
    for (int p = 0; p < np; ++p)
        if ((proxyInfo[p] & 0xffff) == fx && proxyFat[p].x <= proxyFat[p].z) ++children;
    for (int child = 0, left = children; left > 0 && child < 65536; ++child) {
        for (int p = 0; p < np; ++p) {
            const int pi = proxyInfo[p];
            if ((pi & 0xffff) != fx || ((pi >> 16) & 0xffff) != child || proxyFat[p].x > proxyFat[p].z) continue;
            proxyFat[p] = make_float4(kMaxFloat, kMaxFloat, -kMaxFloat, -kMaxFloat);
            proxyMoved[p] = 0;
            proxyFree[nfree++] = p;
            --left;
            break;
        }
    }
    A.proxyFreeCount[ws] = nfree;
}

// b2Fixture::CreateProxies (b2Fixture.cpp:122-139): per child the shape's AABB under xf, fattened by b2_aabbExtension
// (b2DynamicTree::CreateProxy), the most recently freed proxy id or a fresh one, a buffered move.
__device__ __noinline__ void CreateFixtureProxies(const Arena& A, int tile, int lane, int fx, Xf xf) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    Col<int> proxyFree = col(A.proxyFree, tile, A.NP, lane);
    Col<float2> chainVerts = col(A.chainVerts, tile, A.NV, lane);
    const int fi = col(A.fixtureInfo, tile, A.NF, lane)[fx];
    const int shapeType = (fi >> 16) & 3;
    const float* sw = A.shapeWords + ((size_t)tile * A.NS + col(A.fixtureShape, tile, A.NF, lane)[fx]) * kTile + lane;
    // b2Shape::GetChildCount: a chain has count - 1 children (its table entry holds the vertex count), everything else one
    const int children = shapeType == B2G_SHAPE_CHAIN ? __float_as_int(sw[2 * kTile]) - 1 : 1;
    int np = A.proxyCount[ws], nfree = A.proxyFreeCount[ws];
    for (int child = 0; child < children; ++child) {
        if (nfree == 0 && np >= A.NP) { A.status[ws] |= B2GPU_STATUS_SPAWN_OVERFLOW; break; }
        const int slot = nfree > 0 ? proxyFree[--nfree] : np++;
        Vec2 lo, hi;
        ShapeAABB(sw, shapeType, chainVerts, child, xf, &lo, &hi);
        const Vec2 r = V2(kAabbExtension, kAabbExtension);
        lo = lo - r;
        hi = hi + r;
        proxyFat[slot] = make_float4(lo.x, lo.y, hi.x, hi.y);
        proxyInfo[slot] = fx | (child << 16);
        proxyMoved[slot] = 1;
    }
    A.proxyCount[ws] = np;
    A.proxyFreeCount[ws] = nfree;
    A.pendingMoves[ws] = 1;
}

// b2BroadPhase::TouchProxy = BufferMove for every proxy of a fixture
__device__ __noinline__ void TouchFixtureProxies(const Arena& A, int tile, int lane, int fx) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> proxyMoved = col(A.proxyMoved, tile, A.NP, lane);
    const int np = A.proxyCount[ws];
    for (int p = 0; p < np; ++p) {
        if ((proxyInfo[p] & 0xffff) != fx || proxyFat[p].x > proxyFat[p].z) continue;
        proxyMoved[p] = 1;
        A.pendingMoves[ws] = 1;
    }
}

// b2Body::SynchronizeFixtures (b2Body.cpp:442-453) outside a step
__device__ __noinline__ void SynchronizeBodyFixtures(const Arena& A, int tile, int lane, int b) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    const int np = A.proxyCount[ws];
    for (int p = 0; p < np; ++p) {
        const int pi = proxyInfo[p];
        const int fi = fixtureInfo[pi & 0xffff];
        if ((fi & 0xffff) != b || (fi & kFixtureDestroyed)) continue;
        if (SynchronizeProxy(A, tile, lane, p, pi & 0xffff, (pi >> 16) & 0xffff, fi, b)) A.pendingMoves[ws] = 1;
    }
}

// The tail both b2Body::ResetMassData and SetMassData share: the centre of mass moves, the velocity follows it
__device__ __forceinline__ void MoveCenterOfMass(const Arena& A, int tile, int lane, int b, float mass, float invMass, float invI,
                                                 Vec2 localCenter) {
    col(A.bodyDamp, tile, A.NB, lane)[b].w = mass;  // m_mass
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane), bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    const Xf xf = LoadXf(col(A.bodyXf, tile, A.NB, lane), b);
    col(A.bodyMass, tile, A.NB, lane)[b] = make_float4(invMass, invI, localCenter.x, localCenter.y);
    float4 sw = bodySweep[b];
    const Vec2 oldCenter = V2(sw.x, sw.y);
    const Vec2 c = Mul(xf, localCenter);
    sw.x = c.x; sw.y = c.y;
    bodySweep[b] = sw;
    float4 s0 = bodySweep0[b];
    s0.x = c.x; s0.y = c.y;
    bodySweep0[b] = s0;
    float4 v = bodyVel[b];
    const Vec2 lv = V2(v.x, v.y) + Cross(v.z, c - oldCenter);
    v.x = lv.x; v.y = lv.y;
    bodyVel[b] = v;
}

// b2Body::ResetMassData — Dynamics/b2Body.cpp:283-352: the fixtures' mass data summed in fixture-list order (newest first)
__device__ __noinline__ void ResetMassData(const Arena& A, int tile, int lane, int b) {
    const size_t ws = (size_t)tile * kTile + lane;
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane), bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    const int f = col(A.bodyFlags, tile, A.NB, lane)[b];
    const int type = f & kBodyTypeMask;
    if (type == B2G_STATIC_BODY || type == B2G_KINEMATIC_BODY) {
        const float4 xf = col(A.bodyXf, tile, A.NB, lane)[b];
        col(A.bodyMass, tile, A.NB, lane)[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        col(A.bodyDamp, tile, A.NB, lane)[b].w = 0.0f;  // m_mass
        float4 sw = bodySweep[b];
        sw.x = xf.x; sw.y = xf.y;
        bodySweep[b] = sw;
        bodySweep0[b] = make_float4(xf.x, xf.y, sw.z, bodySweep0[b].w);
        return;
    }
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    Col<float4> fixtureMass = col(A.fixtureMass, tile, A.NF, lane);
    float mass = 0.0f, I = 0.0f;
    Vec2 localCenter = V2(0.0f, 0.0f);
    for (int fx = A.fixtureCount[ws] - 1; fx >= 0; --fx) {
        const int fi = fixtureInfo[fx];
        if ((fi & 0xffff) != b || (fi & kFixtureDestroyed)) continue;
        const float4 md = fixtureMass[fx];
        if ((fi & kFixtureHasDensity) == 0) continue;  // if (f->m_density == 0.0f) continue;
        mass += md.x;
        localCenter = localCenter + md.x * V2(md.y, md.z);
        I += md.w;
    }
    float invMass, invI;
    if (mass > 0.0f) {
        invMass = Div(1.0f, mass);
        localCenter = invMass * localCenter;
    } else {
        mass = 1.0f;
        invMass = 1.0f;
    }
    if (I > 0.0f && (f & kBodyFixedRotation) == 0) {
        I -= mass * Dot(localCenter, localCenter);
        invI = Div(1.0f, I);
    } else {
        invI = 0.0f;
    }
    MoveCenterOfMass(A, tile, lane, b, mass, invMass, invI, localCenter);
}

// b2Body::CreateFixture on an existing body (b2Body.cpp:168-214), one call per fixture of the record, in record order: the
// fixture becomes the head of the body's fixture list, gets its proxies under the body's CURRENT transform if the body is
// active, the mass data is reset if it has density, e_newFixture is raised.  Record layout: k_spawn.
__device__ __noinline__ void AttachFixtures(const Arena& A, int tile, int lane, int b, const uint32_t* r, int rf, int rs, int rv) {
    const size_t ws = (size_t)tile * kTile + lane;
    const int rp = (int)r[2];
    const int nf = A.fixtureCount[ws], ns = A.shapeWordCount[ws], nv = A.chainVertCount[ws];
    const uint32_t* p = r + 6 + 25;                       // the body part of the record is not used
    const uint32_t* shapeWords = p + (size_t)rf * 10 + (size_t)rp * 5;
    float* sw = A.shapeWords + ((size_t)tile * A.NS + ns) * kTile + lane;
    for (int k = 0; k < rs; ++k) sw[(size_t)k * kTile] = __uint_as_float(shapeWords[k]);
    Col<float2> cv = col(A.chainVerts, tile, A.NV, lane);
    for (int k = 0; k < rv; ++k) cv[nv + k] = make_float2(__uint_as_float(shapeWords[rs + 2 * k]), __uint_as_float(shapeWords[rs + 2 * k + 1]));
    A.shapeWordCount[ws] = ns + rs;
    A.chainVertCount[ws] = nv + rv;
    const int bf = col(A.bodyFlags, tile, A.NB, lane)[b];
    const Xf xf = LoadXf(col(A.bodyXf, tile, A.NB, lane), b);
    for (int f = 0; f < rf; ++f, p += 10) {
        const int info = (int)p[0], shapeOff = (int)p[1] + ns, fx = nf + f;
        col(A.fixtureInfo, tile, A.NF, lane)[fx] = (info & ~0xffff) | b;
        col(A.fixtureShape, tile, A.NF, lane)[fx] = shapeOff;
        col(A.fixtureMat, tile, A.NF, lane)[fx] = make_float2(__uint_as_float(p[2]), __uint_as_float(p[3]));
        col(A.fixtureFilter, tile, A.NF, lane)[fx] = make_int2((int)p[4], (int)p[5]);
        col(A.fixtureMass, tile, A.NF, lane)[fx] = make_float4(__uint_as_float(p[6]), __uint_as_float(p[7]), __uint_as_float(p[8]), __uint_as_float(p[9]));
        if (((info >> 16) & 3) == B2G_SHAPE_CHAIN) {  // the table entry holds its first vertex relative to the record
            float* e = A.shapeWords + ((size_t)tile * A.NS + shapeOff + 1) * kTile + lane;
            *e = __int_as_float(__float_as_int(*e) + nv);
        }
        A.fixtureCount[ws] = fx + 1;
        if (bf & kBodyActive) CreateFixtureProxies(A, tile, lane, fx, xf);
        if (info & kFixtureHasDensity) ResetMassData(A, tile, lane, b);
    }
    A.newFixture[ws] = 1;
    A.topoDirty[ws] = 1;
}

// b2World::CreateJoint on a running world (b2World.cpp:205-265).  21 words per record: world, then the joint's five
// 4-word columns (the id bits of word 1 are filled in here).  Records are grouped by world, applied in order.
__global__ void k_create_joints(Arena A, const uint32_t* __restrict__ words, int count, int* __restrict__ outId) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int w = (int)words[(size_t)i * 21];
    if (i > 0 && (int)words[(size_t)(i - 1) * 21] == w) return;
    const int tile = w >> 5, lane = w & 31;
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    for (; i < count && (int)words[(size_t)i * 21] == w; ++i) {
        const uint32_t* r = words + (size_t)i * 21 + 1;
        const int nj = A.jointCount[ws], nb = A.bodyCount[ws];
        const int bA = (int)(r[1] & 0xffff), bB = (int)(r[1] >> 16);
        if (!A.hasJoints || nj >= A.NJ) { A.status[ws] |= B2GPU_STATUS_SPAWN_OVERFLOW; outId[i] = -1; continue; }
        if (bA >= nb || bB >= nb || bA == bB || (bodyFlags[bA] & kBodyDestroyed) || (bodyFlags[bB] & kBodyDestroyed)) {
            A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
            outId[i] = -1;
            continue;
        }
        const int id = A.jointSeq[ws];
        auto f4 = [&](float4* base, int k) {
            col(base, tile, A.NJ, lane)[nj] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
        };
        col(A.jointI, tile, A.NJ, lane)[nj] = make_int4((int)((r[0] & 0xffff) | ((uint32_t)id << 16)), (int)r[1], (int)r[2], (int)r[3]);
        f4(A.jointF0, 1); f4(A.jointF1, 2); f4(A.jointF2, 3); f4(A.jointImp, 4);
        A.jointCount[ws] = nj + 1;
        A.jointSeq[ws] = id + 1;
        A.topoDirty[ws] = 1;
        // if the joint prevents collisions, then flag any contacts for filtering (b2World.cpp:255-263)
        if ((r[0] & (1u << 8)) == 0) FlagContactsBetween(A, tile, lane, bA, bB);
        outId[i] = id;
    }
}

// b2World::CreateBody + b2Body::CreateFixture after the first step, from records evaluated by the host layer
// (b2g_host.cpp b2gpu_batch_create_bodies).  Records are grouped by world; the thread that owns a world's first record
// applies that world's records in order.
__global__ void k_spawn(Arena A, const uint32_t* __restrict__ words, const long long* __restrict__ offsets, int count,
                        int* __restrict__ outIndex) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int w = (int)words[offsets[i]];
    if (i > 0 && (int)words[offsets[i - 1]] == w) return;
    const int tile = w >> 5, lane = w & 31;
    const size_t ws = (size_t)tile * kTile + lane;
    for (; i < count && (int)words[offsets[i]] == w; ++i) {
        const uint32_t* r = words + offsets[i];
        const int rf = (int)r[1], rp = (int)r[2], rs = (int)r[3], rv = (int)r[4];
        const int target = (int)r[5];  // -1: a new body; else b2Body::CreateFixture on this existing body
        const int nb = A.bodyCount[ws], nf = A.fixtureCount[ws], np = A.proxyCount[ws], ns = A.shapeWordCount[ws];
        const int nv = A.chainVertCount[ws];
        int nfree = A.proxyFreeCount[ws];
        const int fresh = rp > nfree ? rp - nfree : 0;
        if (nb + (target < 0 ? 1 : 0) > A.NB || nf + rf > A.NF || np + fresh > A.NP || ns + rs > A.NS || nv + rv > A.NV) {
            A.status[ws] |= B2GPU_STATUS_SPAWN_OVERFLOW;
            outIndex[i] = -1;
            continue;
        }
        if (target >= 0) {
            if (target >= nb || (col(A.bodyFlags, tile, A.NB, lane)[target] & kBodyDestroyed)) {
                A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
                outIndex[i] = -1;
                continue;
            }
            AttachFixtures(A, tile, lane, target, r, rf, rs, rv);
            outIndex[i] = nf;  // index of the first new fixture
            continue;
        }
        const uint32_t* p = r + 6;
        auto f4 = [&](float4* base, int cap, int item) {
            col(base, tile, cap, lane)[item] = make_float4(__uint_as_float(p[0]), __uint_as_float(p[1]), __uint_as_float(p[2]), __uint_as_float(p[3]));
            p += 4;
        };
        const int b = nb;  // new highest index = head of m_bodyList
        f4(A.bodyXf, A.NB, b); f4(A.bodySweep, A.NB, b); f4(A.bodySweep0, A.NB, b); f4(A.bodyVel, A.NB, b);
        f4(A.bodyMass, A.NB, b); f4(A.bodyDamp, A.NB, b);
        col(A.bodyForce, tile, A.NB, lane)[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        col(A.bodyFlags, tile, A.NB, lane)[b] = (int)*p++;
        col(A.bodyIsland, tile, A.NB, lane)[b] = -1;
        for (int f = 0; f < rf; ++f) {
            const int info = (int)p[0], shapeOff = (int)p[1] + ns;
            col(A.fixtureInfo, tile, A.NF, lane)[nf + f] = (info & ~0xffff) | b;
            col(A.fixtureShape, tile, A.NF, lane)[nf + f] = shapeOff;
            col(A.fixtureMat, tile, A.NF, lane)[nf + f] = make_float2(__uint_as_float(p[2]), __uint_as_float(p[3]));
            col(A.fixtureFilter, tile, A.NF, lane)[nf + f] = make_int2((int)p[4], (int)p[5]);
            col(A.fixtureMass, tile, A.NF, lane)[nf + f] =
                make_float4(__uint_as_float(p[6]), __uint_as_float(p[7]), __uint_as_float(p[8]), __uint_as_float(p[9]));
            p += 10;
        }
        Col<int> proxyFree = col(A.proxyFree, tile, A.NP, lane);
        int npNow = np;
        for (int q = 0; q < rp; ++q) {  // b2Fixture::CreateProxies -> b2DynamicTree::AllocateNode: most recently freed id first
            const int slot = nfree > 0 ? proxyFree[--nfree] : npNow++;
            col(A.proxyFat, tile, A.NP, lane)[slot] = make_float4(__uint_as_float(p[0]), __uint_as_float(p[1]), __uint_as_float(p[2]), __uint_as_float(p[3]));
            const int info = (int)p[4];
            col(A.proxyInfo, tile, A.NP, lane)[slot] = (nf + (info & 0xffff)) | (info & ~0xffff);
            col(A.proxyMoved, tile, A.NP, lane)[slot] = 1;  // b2BroadPhase::CreateProxy buffers a move
            p += 5;
        }
        float* sw = A.shapeWords + ((size_t)tile * A.NS + ns) * kTile + lane;
        for (int k = 0; k < rs; ++k) sw[(size_t)k * kTile] = __uint_as_float(p[k]);
        p += rs;
        Col<float2> cv = col(A.chainVerts, tile, A.NV, lane);
        for (int k = 0; k < rv; ++k) cv[nv + k] = make_float2(__uint_as_float(p[2 * k]), __uint_as_float(p[2 * k + 1]));
        // chain entries of the shape table hold their first vertex index: rebase it
        for (int f = 0; f < rf; ++f) {
            const int fi = col(A.fixtureInfo, tile, A.NF, lane)[nf + f];
            if (((fi >> 16) & 3) != B2G_SHAPE_CHAIN) continue;
            float* e = A.shapeWords + ((size_t)tile * A.NS + col(A.fixtureShape, tile, A.NF, lane)[nf + f] + 1) * kTile + lane;
            *e = __int_as_float(__float_as_int(*e) + nv);  // (two fixtures never share a chain entry: its first vertex is unique)
        }
        A.bodyCount[ws] = nb + 1;
        A.fixtureCount[ws] = nf + rf;
        A.proxyCount[ws] = npNow;
        A.proxyFreeCount[ws] = nfree;
        A.shapeWordCount[ws] = ns + rs;
        A.chainVertCount[ws] = nv + rv;
        A.newFixture[ws] = 1;
        A.topoDirty[ws] = 1;
        outIndex[i] = b;
    }
}

// Compaction of the per-world event arenas: world w's events go to out[offsets[w] ...] (b2gpu_batch_read_events).
__global__ void k_gather_events(Arena A, const long long* __restrict__ offsets, b2gpu_event* __restrict__ out, long long capacity) {
    const int w = blockIdx.x;  // one CTA per world, threads stride over its events
    if (w >= A.numWorlds) return;
    const int tile = w >> 5, lane = w & 31;
    const long long o = offsets[w];
    const int n = (int)(offsets[w + 1] - o);
    Col<int4> ea = col(A.eventA, tile, A.NE, lane);
    Col<float4> eb = col(A.eventB, tile, A.NE, lane);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (o + i >= capacity) break;
        const int4 a = ea[i];
        const float4 f = eb[i];
        b2gpu_event e;
        e.type = a.x;
        e.fixtureA = a.y & 0xffff; e.childA = (a.y >> 16) & 0xffff;
        e.fixtureB = a.z & 0xffff; e.childB = (a.z >> 16) & 0xffff;
        e.count = a.w;
        e.normalImpulses[0] = f.x; e.normalImpulses[1] = f.y;
        e.tangentImpulses[0] = f.z; e.tangentImpulses[1] = f.w;
        out[o + i] = e;
    }
}


// Between-step body edits (include/b2gpu.h B2GPU_CMD_*).  The command array is grouped by world; the thread that
// owns the first command of a world applies that world's commands in order, exactly as the reference's user
// would have made the calls (Box2D/Dynamics/b2Body.h:738-833, 408-467, 634-653; b2Body.cpp:418-440).
__global__ void k_body_commands(Arena A, const b2gpu_body_command* __restrict__ cmds, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int w = cmds[i].world;
    if (i > 0 && cmds[i - 1].world == w) return;  // not the first command of its world
    const int tile = w >> 5, lane = w & 31;
    const size_t ws = (size_t)tile * kTile + lane;
    Col<int> bodyFlags = col(A.bodyFlags, tile, A.NB, lane);
    Col<float4> bodySweep = col(A.bodySweep, tile, A.NB, lane);
    Col<float4> bodySweep0 = col(A.bodySweep0, tile, A.NB, lane);
    Col<float4> bodyVel = col(A.bodyVel, tile, A.NB, lane);
    Col<float4> bodyForce = col(A.bodyForce, tile, A.NB, lane);
    Col<float4> bodyMass = col(A.bodyMass, tile, A.NB, lane);
    Col<float4> bodyXf = col(A.bodyXf, tile, A.NB, lane);
    const int nb = A.bodyCount[ws];
    for (; i < count && cmds[i].world == w; ++i) {
        const b2gpu_body_command c = cmds[i];
        if (c.op == B2GPU_CMD_DESTROY_JOINT) {  // b2World::DestroyJoint — Dynamics/b2World.cpp:267-360
            int found = -1;
            Col<int4> jI = col(A.jointI, tile, A.NJ, lane);
            const int nj = A.hasJoints ? A.jointCount[ws] : 0;
            for (int j = 0; j < nj && found < 0; ++j)
                if (((jI[j].x >> 16) & 0xffff) == c.wake) found = j;
            if (found < 0) { A.status[ws] |= B2GPU_STATUS_BAD_COMMAND; continue; }
            const int4 ji = jI[found];
            const int jA = ji.y & 0xffff, jB = (ji.y >> 16) & 0xffff;
            if (WakeBody(bodyFlags, bodySweep, jA)) A.topoDirty[ws] = 1;  // bodyA->SetAwake(true); bodyB->SetAwake(true)
            if (WakeBody(bodyFlags, bodySweep, jB)) A.topoDirty[ws] = 1;
            Col<float4> jF0 = col(A.jointF0, tile, A.NJ, lane), jF1 = col(A.jointF1, tile, A.NJ, lane);
            Col<float4> jF2 = col(A.jointF2, tile, A.NJ, lane), jImp = col(A.jointImp, tile, A.NJ, lane);
            for (int j = found; j + 1 < nj; ++j) { jI[j] = jI[j + 1]; jF0[j] = jF0[j + 1]; jF1[j] = jF1[j + 1]; jF2[j] = jF2[j + 1]; jImp[j] = jImp[j + 1]; }
            A.jointCount[ws] = nj - 1;
            A.topoDirty[ws] = 1;
            if ((ji.x & (1 << 8)) == 0) FlagContactsBetween(A, tile, lane, jA, jB);  // collideConnected == false: the pair may collide now
            continue;
        }
        if (c.body < 0 || c.body >= nb || c.op < 0 || c.op >= B2GPU_CMD_COUNT) {
            A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
            continue;
        }
        const int b = c.body;
        int f = bodyFlags[b];
        if (f & kBodyDestroyed) {  // the reference would be handed a dangling b2Body*
            A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
            continue;
        }
        const int type = f & kBodyTypeMask;
        auto wake = [&]() {  // SetAwake(true)
            if (WakeBody(bodyFlags, bodySweep, b)) A.topoDirty[ws] = 1;
            f |= kBodyAwake;
        };
        switch (c.op) {
            case B2GPU_CMD_APPLY_FORCE:
            case B2GPU_CMD_APPLY_FORCE_TO_CENTER:
            case B2GPU_CMD_APPLY_TORQUE: {
                if (type != B2G_DYNAMIC_BODY) break;
                if (c.wake && (f & kBodyAwake) == 0) wake();
                if ((f & kBodyAwake) == 0) break;
                float4 cur = bodyForce[b];
                if (c.op == B2GPU_CMD_APPLY_TORQUE) {
                    cur.z += c.v[0];
                } else {
                    cur.x += c.v[0];
                    cur.y += c.v[1];
                    if (c.op == B2GPU_CMD_APPLY_FORCE) {
                        const float4 sw = bodySweep[b];
                        cur.z += Cross(V2(c.p[0], c.p[1]) - V2(sw.x, sw.y), V2(c.v[0], c.v[1]));
                    }
                }
                bodyForce[b] = cur;
            } break;
            case B2GPU_CMD_APPLY_LINEAR_IMPULSE:
            case B2GPU_CMD_APPLY_ANGULAR_IMPULSE: {
                if (type != B2G_DYNAMIC_BODY) break;
                if (c.wake && (f & kBodyAwake) == 0) wake();
                if ((f & kBodyAwake) == 0) break;
                const float4 m = bodyMass[b];
                float4 v = bodyVel[b];
                if (c.op == B2GPU_CMD_APPLY_LINEAR_IMPULSE) {
                    const float4 sw = bodySweep[b];
                    const Vec2 imp = V2(c.v[0], c.v[1]);
                    const Vec2 lv = V2(v.x, v.y) + m.x * imp;
                    v.x = lv.x;
                    v.y = lv.y;
                    v.z += m.y * Cross(V2(c.p[0], c.p[1]) - V2(sw.x, sw.y), imp);
                } else {
                    v.z += m.y * c.v[0];
                }
                bodyVel[b] = v;
            } break;
            case B2GPU_CMD_SET_LINEAR_VELOCITY: {
                if (type == B2G_STATIC_BODY) break;
                if (Dot(V2(c.v[0], c.v[1]), V2(c.v[0], c.v[1])) > 0.0f) wake();
                float4 v = bodyVel[b];
                v.x = c.v[0];
                v.y = c.v[1];
                bodyVel[b] = v;
            } break;
            case B2GPU_CMD_SET_ANGULAR_VELOCITY: {
                if (type == B2G_STATIC_BODY) break;
                if (c.v[0] * c.v[0] > 0.0f) wake();
                float4 v = bodyVel[b];
                v.z = c.v[0];
                bodyVel[b] = v;
            } break;
            case B2GPU_CMD_MOUSE_SET_TARGET: {  // b2MouseJoint::SetTarget — Dynamics/Joints/b2MouseJoint.cpp:54-61
                int found = -1;
                if (A.hasJoints) {
                    Col<int4> jI = col(A.jointI, tile, A.NJ, lane);
                    for (int j = A.jointCount[ws] - 1; j >= 0 && found < 0; --j) {  // the body's joint list is newest first
                        const int4 ji = jI[j];
                        if ((ji.x & 0xff) == B2G_JOINT_MOUSE && ((ji.y >> 16) & 0xffff) == b) found = j;
                    }
                }
                if (found < 0) { A.status[ws] |= B2GPU_STATUS_BAD_COMMAND; break; }
                if ((f & kBodyAwake) == 0) wake();
                Col<float4> jF0 = col(A.jointF0, tile, A.NJ, lane);
                float4 t = jF0[found];
                t.x = c.v[0];
                t.y = c.v[1];
                jF0[found] = t;
            } break;
            case B2GPU_CMD_DESTROY_BODY: {  // b2World::DestroyBody — Dynamics/b2World.cpp:124-203
                A.topoDirty[ws] = 1;
                // 1. the attached joints, in the body's joint-list order (newest first): b2World::DestroyJoint wakes
                //    both bodies (b2World.cpp:302-304); then a stable erase keeps m_jointList's order
                if (A.hasJoints) {
                    Col<int4> jI = col(A.jointI, tile, A.NJ, lane);
                    Col<float4> jF0 = col(A.jointF0, tile, A.NJ, lane), jF1 = col(A.jointF1, tile, A.NJ, lane);
                    Col<float4> jF2 = col(A.jointF2, tile, A.NJ, lane), jImp = col(A.jointImp, tile, A.NJ, lane);
                    const int nj = A.jointCount[ws];
                    bool any = false;
                    for (int j = nj - 1; j >= 0; --j) {
                        const int4 ji = jI[j];
                        const int jA = ji.y & 0xffff, jB = (ji.y >> 16) & 0xffff;
                        if (jA != b && jB != b) continue;
                        WakeBody(bodyFlags, bodySweep, jA);
                        WakeBody(bodyFlags, bodySweep, jB);
                        any = true;
                    }
                    if (any) {
                        int wpos = 0;
                        for (int j = 0; j < nj; ++j) {
                            const int4 ji = jI[j];
                            if ((ji.y & 0xffff) == b || ((ji.y >> 16) & 0xffff) == b) continue;
                            if (wpos != j) { jI[wpos] = ji; jF0[wpos] = jF0[j]; jF1[wpos] = jF1[j]; jF2[wpos] = jF2[j]; jImp[wpos] = jImp[j]; }
                            ++wpos;
                        }
                        A.jointCount[ws] = wpos;
                    }
                }
                // 2. the attached contacts, in the body's contact-list order (newest first)
                DestroyContactsOf(A, tile, lane, b, -1);
                // 3. the fixtures' proxies leave the broad phase (b2Fixture::DestroyProxies), fixture list order = newest first
                {
                    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                    for (int fx = A.fixtureCount[ws] - 1; fx >= 0; --fx) {
                        const int fi = fixtureInfo[fx];
                        if ((fi & 0xffff) != b || (fi & kFixtureDestroyed)) continue;
                        fixtureInfo[fx] = fi | kFixtureDestroyed;
                        DestroyFixtureProxies(A, tile, lane, fx);
                    }
                }
                // 4. the body itself: a tombstone that is never a seed, never woken, never touched again
                bodyFlags[b] = (bodyFlags[b] & kBodyTypeMask) | kBodyDestroyed;
                bodyVel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            } break;
            case B2GPU_CMD_SET_FILTER: {  // b2Fixture::SetFilterData + Refilter — Dynamics/b2Fixture.cpp:185-215
                const int fx = c.wake;
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                if (fx < 0 || fx >= A.fixtureCount[ws] || (fixtureInfo[fx] & 0xffff) != b || (fixtureInfo[fx] & kFixtureDestroyed)) {
                    A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
                    break;
                }
                col(A.fixtureFilter, tile, A.NF, lane)[fx] = make_int2(__float_as_int(c.v[0]), __float_as_int(c.v[1]));
                // flag the fixture's contacts for filtering
                Col<int4> cA = col(A.contactA, tile, A.NC, lane);
                Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
                const int nc = A.contactCount[ws];
                for (int k = 0; k < nc; ++k) {
                    int4 a = cA[k];
                    if ((proxyInfo[a.x & 0xffff] & 0xffff) != fx && (proxyInfo[(a.x >> 16) & 0xffff] & 0xffff) != fx) continue;
                    a.z |= kContactFilter;
                    cA[k] = a;
                }
                if (f & kBodyActive) TouchFixtureProxies(A, tile, lane, fx);  // an inactive body's fixtures have no proxies
            } break;
            case B2GPU_CMD_SET_TYPE: {  // b2Body::SetType — Dynamics/b2Body.cpp:115-166
                const int newType = c.wake;
                if (newType < B2G_STATIC_BODY || newType > B2G_DYNAMIC_BODY) { A.status[ws] |= B2GPU_STATUS_BAD_COMMAND; break; }
                if (newType == type) break;
                f = (f & ~kBodyTypeMask) | newType;
                bodyFlags[b] = f;
                A.topoDirty[ws] = 1;
                ResetMassData(A, tile, lane, b);
                if (newType == B2G_STATIC_BODY) {
                    bodyVel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    const float4 sw = bodySweep[b];
                    bodySweep0[b] = make_float4(sw.x, sw.y, sw.z, 0.0f);
                    if (f & kBodyActive) SynchronizeBodyFixtures(A, tile, lane, b);
                }
                wake();
                bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                DestroyContactsOf(A, tile, lane, b, -1);
                if (f & kBodyActive) {  // touch the proxies so that new contacts will be created (when appropriate)
                    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                    for (int fx = A.fixtureCount[ws] - 1; fx >= 0; --fx)
                        if ((fixtureInfo[fx] & 0xffff) == b && (fixtureInfo[fx] & kFixtureDestroyed) == 0) TouchFixtureProxies(A, tile, lane, fx);
                }
            } break;
            case B2GPU_CMD_SET_ACTIVE: {  // b2Body::SetActive — Dynamics/b2Body.cpp:455-510
                const bool flag = c.wake != 0;
                if (flag == ((f & kBodyActive) != 0)) break;
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                A.topoDirty[ws] = 1;
                if (flag) {
                    bodyFlags[b] = f | kBodyActive;
                    const Xf xf = LoadXf(bodyXf, b);
                    for (int fx = A.fixtureCount[ws] - 1; fx >= 0; --fx)  // m_fixtureList: newest first
                        if ((fixtureInfo[fx] & 0xffff) == b && (fixtureInfo[fx] & kFixtureDestroyed) == 0) CreateFixtureProxies(A, tile, lane, fx, xf);
                } else {
                    bodyFlags[b] = f & ~kBodyActive;
                    for (int fx = A.fixtureCount[ws] - 1; fx >= 0; --fx)
                        if ((fixtureInfo[fx] & 0xffff) == b && (fixtureInfo[fx] & kFixtureDestroyed) == 0) DestroyFixtureProxies(A, tile, lane, fx);
                    DestroyContactsOf(A, tile, lane, b, -1);
                }
            } break;
            case B2GPU_CMD_SET_MASS_DATA: {  // b2Body::SetMassData — Dynamics/b2Body.cpp:354-394
                if (type != B2G_DYNAMIC_BODY) break;
                float mass = c.v[0];
                if (mass <= 0.0f) mass = 1.0f;
                const float invMass = Div(1.0f, mass);
                float invI = 0.0f;
                const Vec2 center = V2(c.p[0], c.p[1]);
                if (c.v[1] > 0.0f && (f & kBodyFixedRotation) == 0) {
                    const float I = c.v[1] - mass * Dot(center, center);
                    invI = Div(1.0f, I);
                }
                MoveCenterOfMass(A, tile, lane, b, mass, invMass, invI, center);
            } break;
            case B2GPU_CMD_RESET_MASS_DATA: {
                ResetMassData(A, tile, lane, b);
            } break;
            case B2GPU_CMD_DESTROY_FIXTURE: {  // b2Body::DestroyFixture — Dynamics/b2Body.cpp:216-281
                const int fx = c.wake;
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                if (fx < 0 || fx >= A.fixtureCount[ws] || (fixtureInfo[fx] & 0xffff) != b || (fixtureInfo[fx] & kFixtureDestroyed)) {
                    A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
                    break;
                }
                DestroyContactsOf(A, tile, lane, b, fx);
                if (f & kBodyActive) DestroyFixtureProxies(A, tile, lane, fx);
                fixtureInfo[fx] |= kFixtureDestroyed;
                ResetMassData(A, tile, lane, b);
            } break;
            case B2GPU_CMD_SET_DAMPING: {
                Col<float4> bodyDamp = col(A.bodyDamp, tile, A.NB, lane);
                float4 d = bodyDamp[b];
                d.x = c.v[0];
                d.y = c.v[1];
                bodyDamp[b] = d;
            } break;
            case B2GPU_CMD_SET_GRAVITY_SCALE: {
                Col<float4> bodyDamp = col(A.bodyDamp, tile, A.NB, lane);
                float4 d = bodyDamp[b];
                d.z = c.v[0];
                bodyDamp[b] = d;
            } break;
            case B2GPU_CMD_SET_BULLET: {
                bodyFlags[b] = c.wake ? (f | kBodyBullet) : (f & ~kBodyBullet);
            } break;
            case B2GPU_CMD_SET_SLEEPING_ALLOWED: {  // b2Body.h:655-667
                if (c.wake) {
                    bodyFlags[b] = f | kBodyAutoSleep;
                } else {
                    bodyFlags[b] = f & ~kBodyAutoSleep;
                    f &= ~kBodyAutoSleep;
                    wake();
                }
            } break;
            case B2GPU_CMD_SET_FIXED_ROTATION: {  // b2Body.cpp:512-534
                const bool status = (f & kBodyFixedRotation) != 0, flag = c.wake != 0;
                if (status == flag) break;
                bodyFlags[b] = flag ? (f | kBodyFixedRotation) : (f & ~kBodyFixedRotation);
                float4 v = bodyVel[b];
                v.z = 0.0f;
                bodyVel[b] = v;
                ResetMassData(A, tile, lane, b);
            } break;
            case B2GPU_CMD_SET_DENSITY: {  // b2Fixture::SetDensity (b2Fixture.h:313-317): the body's mass waits for ResetMassData
                const int fx = c.wake;
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                if (fx < 0 || fx >= A.fixtureCount[ws] || (fixtureInfo[fx] & 0xffff) != b || (fixtureInfo[fx] & kFixtureDestroyed)) {
                    A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
                    break;
                }
                col(A.fixtureMass, tile, A.NF, lane)[fx] = make_float4(c.v[0], c.p[0], c.p[1], c.v[1]);
                const bool dense = c.v[0] != 0.0f || c.v[1] != 0.0f;
                fixtureInfo[fx] = dense ? (fixtureInfo[fx] | kFixtureHasDensity) : (fixtureInfo[fx] & ~kFixtureHasDensity);
            } break;
            case B2GPU_CMD_SET_SENSOR:
            case B2GPU_CMD_SET_MATERIAL: {
                const int fx = c.wake;
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                if (fx < 0 || fx >= A.fixtureCount[ws] || (fixtureInfo[fx] & 0xffff) != b || (fixtureInfo[fx] & kFixtureDestroyed)) {
                    A.status[ws] |= B2GPU_STATUS_BAD_COMMAND;
                    break;
                }
                if (c.op == B2GPU_CMD_SET_MATERIAL) {
                    col(A.fixtureMat, tile, A.NF, lane)[fx] = make_float2(c.v[0], c.v[1]);
                    break;
                }
                const int fi = fixtureInfo[fx];
                const bool sensor = c.v[0] != 0.0f;
                if (sensor == ((fi >> 18 & 1) != 0)) break;
                wake();  // m_body->SetAwake(true)
                fixtureInfo[fx] = sensor ? (fi | (1 << 18)) : (fi & ~(1 << 18));
                // b2Contact::Update reads both fixtures' sensor flags every time (b2Contact.cpp:181-183): the cached bit of the
                // fixture's contacts follows; island topology may change with it
                Col<int4> cA = col(A.contactA, tile, A.NC, lane);
                Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
                const int nc = A.contactCount[ws];
                for (int k = 0; k < nc; ++k) {
                    int4 a = cA[k];
                    const int fA = proxyInfo[a.x & 0xffff] & 0xffff, fB = proxyInfo[(a.x >> 16) & 0xffff] & 0xffff;
                    if (fA != fx && fB != fx) continue;
                    const bool s2 = ((fixtureInfo[fA] | fixtureInfo[fB]) >> 18 & 1) != 0;
                    a.z = s2 ? (a.z | kContactSensor) : (a.z & ~kContactSensor);
                    cA[k] = a;
                }
                A.topoDirty[ws] = 1;
            } break;
            case B2GPU_CMD_SET_AWAKE: {
                if (c.wake) {
                    wake();
                } else {
                    if (f & kBodyAwake) A.topoDirty[ws] = 1;
                    bodyFlags[b] = f & ~kBodyAwake;
                    float4 sw = bodySweep[b];
                    sw.w = 0.0f;
                    bodySweep[b] = sw;
                    bodyVel[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    bodyForce[b] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
            } break;
            case B2GPU_CMD_SET_TRANSFORM: {
                Xf xf;
                xf.q = MakeRot(c.p[0]);
                xf.p = V2(c.v[0], c.v[1]);
                bodyXf[b] = make_float4(xf.p.x, xf.p.y, xf.q.s, xf.q.c);
                const float4 m = bodyMass[b];
                const Vec2 cen = Mul(xf, V2(m.z, m.w));
                float4 sw = bodySweep[b];
                sw.x = cen.x; sw.y = cen.y; sw.z = c.p[0];
                bodySweep[b] = sw;
                float4 s0 = bodySweep0[b];
                s0.x = cen.x; s0.y = cen.y; s0.z = c.p[0];
                bodySweep0[b] = s0;
                // f->Synchronize(broadPhase, m_xf, m_xf) for every fixture of the body (b2Fixture.cpp:154-177)
                Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
                Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
                Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
                Col<float2> chainVerts = col(A.chainVerts, tile, A.NV, lane);
                const int np = A.proxyCount[ws];
                for (int p = 0; p < np; ++p) {
                    const int pi = proxyInfo[p];
                    const int fixture = pi & 0xffff, child = pi >> 16;
                    const int fi = fixtureInfo[fixture];
                    if ((fi & 0xffff) != b) continue;
                    const float* sw_ = A.shapeWords + ((size_t)tile * A.NS + col(A.fixtureShape, tile, A.NF, lane)[fixture]) * kTile + lane;
                    Vec2 lo1, hi1;
                    ShapeAABB(sw_, (fi >> 16) & 3, chainVerts, child, xf, &lo1, &hi1);
                    Vec2 lo = Min(lo1, lo1), hi = Max(hi1, hi1);
                    const Vec2 displacement = xf.p - xf.p;
                    const float4 fat = proxyFat[p];
                    if (fat.x > fat.z) continue;  // freed slot
                    if (fat.x <= lo.x && fat.y <= lo.y && hi.x <= fat.z && hi.y <= fat.w) continue;
                    const Vec2 r = V2(kAabbExtension, kAabbExtension);
                    lo = lo - r;
                    hi = hi + r;
                    const Vec2 d = kAabbMultiplier * displacement;
                    if (d.x < 0.0f) lo.x += d.x; else hi.x += d.x;
                    if (d.y < 0.0f) lo.y += d.y; else hi.y += d.y;
                    proxyFat[p] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    col(A.proxyMoved, tile, A.NP, lane)[p] = 1;
                    A.pendingMoves[ws] = 1;
                }
            } break;
        }
    }
}

// ---- stateless entry points: the kernels' building blocks on flat inputs (unit-level parity) ----
struct FlatPoly {  // 35 floats: count, 8 vertices, 8 normals, centroid (oracle/ref_harness.cpp fill_poly)
    const float* w;
    __device__ __forceinline__ int count() const { return (int)w[0]; }
    __device__ __forceinline__ float radius() const { return kPolygonRadius; }
    __device__ __forceinline__ Vec2 centroid() const { return V2(w[33], w[34]); }
    __device__ __forceinline__ Vec2 vert(int i) const { return V2(w[1 + 2 * i], w[2 + 2 * i]); }
    __device__ __forceinline__ Vec2 normal(int i) const { return V2(w[17 + 2 * i], w[18 + 2 * i]); }
};

__global__ void k_collide_pairs(int kind, int n, const float* __restrict__ shapeA, const float* __restrict__ xfA,
                                const float* __restrict__ shapeB, const float* __restrict__ xfB, int* __restrict__ ints,
                                float* __restrict__ floats) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Xf xa, xb;
    xa.p = V2(xfA[4 * i], xfA[4 * i + 1]); xa.q.s = xfA[4 * i + 2]; xa.q.c = xfA[4 * i + 3];
    xb.p = V2(xfB[4 * i], xfB[4 * i + 1]); xb.q.s = xfB[4 * i + 2]; xb.q.c = xfB[4 * i + 3];
    Manifold m;
    m.pointCount = 0; m.type = 0; m.key0 = m.key1 = 0;
    m.lp0 = m.lp1 = m.localNormal = m.localPoint = V2(0.0f, 0.0f);
    auto circle = [](const float* c) { CircleData d; d.p = V2(c[0], c[1]); d.radius = c[2]; return d; };
    auto edge = [](const float* e) {
        EdgeData d;
        d.v0 = V2(e[0], e[1]); d.v1 = V2(e[2], e[3]); d.v2 = V2(e[4], e[5]); d.v3 = V2(e[6], e[7]);
        d.has0 = e[8] != 0.0f; d.has3 = e[9] != 0.0f; d.radius = kPolygonRadius;
        return d;
    };
    switch (kind) {
        case B2GPU_COLLIDE_POLYGONS:
            CollidePolygons(&m, FlatPoly{shapeA + 35 * (size_t)i}, xa, FlatPoly{shapeB + 35 * (size_t)i}, xb); break;
        case B2GPU_COLLIDE_CIRCLES:
            CollideCircles(&m, circle(shapeA + 3 * (size_t)i), xa, circle(shapeB + 3 * (size_t)i), xb); break;
        case B2GPU_COLLIDE_POLYGON_CIRCLE:
            CollidePolygonAndCircle(&m, FlatPoly{shapeA + 35 * (size_t)i}, xa, circle(shapeB + 3 * (size_t)i), xb); break;
        case B2GPU_COLLIDE_EDGE_CIRCLE:
            CollideEdgeAndCircle(&m, edge(shapeA + 10 * (size_t)i), xa, circle(shapeB + 3 * (size_t)i), xb); break;
        default:
            CollideEdgeAndPolygon(&m, edge(shapeA + 10 * (size_t)i), xa, FlatPoly{shapeB + 35 * (size_t)i}, xb); break;
    }
    int* I = ints + 4 * (size_t)i;
    float* F = floats + 8 * (size_t)i;
    I[0] = m.pointCount; I[1] = m.type;
    I[2] = m.pointCount > 0 ? (int)m.key0 : 0;
    I[3] = m.pointCount > 1 ? (int)m.key1 : 0;
    for (int k = 0; k < 8; ++k) F[k] = 0.0f;
    if (m.pointCount > 0) {
        F[0] = m.localNormal.x; F[1] = m.localNormal.y; F[2] = m.localPoint.x; F[3] = m.localPoint.y;
        F[4] = m.lp0.x; F[5] = m.lp0.y;
        if (m.pointCount > 1) { F[6] = m.lp1.x; F[7] = m.lp1.y; }
    }
}

__global__ void k_sincosf(const float* __restrict__ x, float* __restrict__ s, float* __restrict__ c, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) SinCos(x[i], &s[i], &c[i]);
}

}  // namespace

constexpr int kGraphChunk = 8;

struct b2gpu_batch {
    b2gpu_batch_desc desc;
    Arena A;
    cudaStream_t stream = nullptr;
    // World groups: the step's kernel chain is enqueued once per group of tiles, group 0 on `stream`, the others on
    // their own streams (forked from and joined back into `stream` with events, inside or outside graph capture).
    // Groups are independent worlds, so one group's sequential kernels (64 worlds per SM, 2 of 4 schedulers) overlap
    // another group's parallel kernels on the same SMs.
    static constexpr int kMaxGroups = 8;
    int groups = 1;
    cudaStream_t groupStream[kMaxGroups] = {nullptr};
    cudaEvent_t evFork = nullptr, evJoin[kMaxGroups] = {nullptr}, evSeq[kMaxGroups] = {nullptr};
    bool pipelineGroups = false;  // B2GPU_PIPELINE=1 orders the groups' sequential parts (measured slower than free-running)
    std::vector<Field> fields;
    std::vector<void*> allocations;
    bool needInitialPairs = true;  // b2World::e_newFixture
    bool needProxyIndex = true;    // the body -> proxies index has to be (re)built before the next step
    bool anyContinuous = false;    // some world has B2G_WORLD_CONTINUOUS: the step ends with k_toi
    // Can a dynamic body of this batch be asleep?  (a world that allows sleeping, a body built or set asleep, a restored
    // snapshot.)  While it cannot, the forces of b2gpu_batch_step_io wake nobody, so they may reach the device behind
    // k_collide / k_islands (which do not read them) and their host copy hides under those kernels.
    bool maySleep = false;
    cudaStream_t copyStream[kMaxGroups] = {nullptr};
    cudaEvent_t evCopy[kMaxGroups] = {nullptr};
    int blockSize = 32;
    bool tileSolver = false;       // b2g_solve.cuh path: the tile's body table fits in shared memory
    size_t solveSmem = 0;
    int islandLanes = 32;          // worlds per CTA of k_islands
    size_t islandSmem = 0;
    int islandGrid = 148;          // CTAs of k_islands' persistent grid
    bool collideSmemOk = false;    // k_collide's result bytes fit in shared memory
    int64_t launches = 0;
    size_t arenaBytes = 0;
    // graph of one step, keyed by its parameters
    cudaGraphExec_t graphExec = nullptr;       // one step
    cudaGraphExec_t graphExecChunk = nullptr;  // kGraphChunk steps (several groups only)
    float graphDt = 0.0f;
    int graphVel = 0, graphPos = 0;
    // graph of one b2gpu_batch_step_io call (kernels and the groups' host copies), keyed by the host pointers too
    cudaGraphExec_t graphExecIo = nullptr;
    const float* ioGraphForces = nullptr;
    float* ioGraphTransforms = nullptr;
    float ioGraphDt = 0.0f;
    int ioGraphVel = 0, ioGraphPos = 0, ioGraphGroups = 0;
    int64_t ioGraphLaunches = 0;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    float lastStepMs = 0.0f;
    // per-step I/O staging
    float* ioDev = nullptr;
    size_t ioBytes = 0;
    // b2gpu_batch_io_select: the bodies (device index lists; null = every body slot) and the format of the dense I/O
    int* ioForceSel = nullptr;
    int* ioPoseSel = nullptr;
    int ioForceCount = 0, ioPoseCount = 0;  // entries per world (NB while nothing is selected)
    int ioPoseFormat = B2GPU_POSE_XF;
    b2gpu_body_command* cmdDev = nullptr;
    size_t cmdBytes = 0;
};

namespace {

template <class T>
int arenaAlloc(b2gpu_batch* b, T** ptr, size_t itemsPerWorld, size_t elemBytes) {
    size_t bytes = (size_t)b->A.numTiles * kTile * itemsPerWorld * elemBytes;
    if (bytes == 0) bytes = 16;
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes));
    CUDA_TRY(cudaMemsetAsync(p, 0, bytes, b->stream));
    b->allocations.push_back(p);
    b->arenaBytes += bytes;
    *ptr = (T*)p;
    return B2GPU_OK;
}

#define ALLOC(ptr, items, T)                                                       \
    do {                                                                           \
        int rc__ = arenaAlloc(b, &b->A.ptr, (size_t)(items), sizeof(T));           \
        if (rc__ != B2GPU_OK) return rc__;                                         \
    } while (0)

int allocateArena(b2gpu_batch* b) {
    Arena& A = b->A;
    const b2gpu_batch_desc& d = b->desc;
    A.numWorlds = d.numWorlds;
    A.numTiles = (d.numWorlds + kTile - 1) / kTile;
    A.NB = d.maxBodies; A.NF = d.maxFixtures; A.NP = d.maxProxies; A.NC = d.maxContacts;
    A.NS = d.maxShapeWords; A.NV = d.maxChainVerts > 0 ? d.maxChainVerts : 1; A.NE = d.maxEvents > 0 ? d.maxEvents : 1;
    A.NJ = d.maxJoints > 0 ? d.maxJoints : 1;
    A.hasJoints = d.maxJoints > 0 ? 1 : 0;
    A.eventMask = d.eventMask;
    A.alwaysRebuildIslands = (d.reserved >> 1) & 1;
    ALLOC(gravity, 1, float2); ALLOC(worldFlags, 1, int); ALLOC(invDt0, 1, float);
    ALLOC(bodyCount, 1, int); ALLOC(proxyCount, 1, int); ALLOC(contactCount, 1, int);
    ALLOC(status, 1, int); ALLOC(eventCount, 1, int); ALLOC(islandCount, 1, int); ALLOC(stepCount, 1, int);
    ALLOC(topoDirty, 1, int); ALLOC(pendingMoves, 1, int);
    ALLOC(fixtureCount, 1, int); ALLOC(shapeWordCount, 1, int); ALLOC(chainVertCount, 1, int); ALLOC(proxyFreeCount, 1, int);
    ALLOC(newFixture, 1, int); ALLOC(proxyFree, A.NP, int);
    CUDA_TRY(cudaMemsetAsync(A.topoDirty, 1, (size_t)A.numTiles * kTile * sizeof(int), b->stream));
    ALLOC(bodyXf, A.NB, float4); ALLOC(bodySweep, A.NB, float4); ALLOC(bodySweep0, A.NB, float4);
    ALLOC(bodyVel, A.NB, float4); ALLOC(bodyForce, A.NB, float4); ALLOC(bodyMass, A.NB, float4);
    ALLOC(bodyDamp, A.NB, float4); ALLOC(bodyFlags, A.NB, int); ALLOC(bodyIsland, A.NB, int);
    ALLOC(fixtureInfo, A.NF, int); ALLOC(fixtureShape, A.NF, int); ALLOC(fixtureMat, A.NF, float2);
    ALLOC(fixtureFilter, A.NF, int2); ALLOC(fixtureMass, A.NF, float4);
    ALLOC(proxyFat, A.NP, float4); ALLOC(proxyInfo, A.NP, int); ALLOC(proxyMoved, A.NP, int);
    ALLOC(shapeWords, A.NS, float); ALLOC(chainVerts, A.NV, float2);
    ALLOC(contactA, A.NC, int4); ALLOC(contactK, A.NC, int4); ALLOC(contactS, A.NC, int2);
    ALLOC(contactP0, A.NC, float4); ALLOC(contactP1, A.NC, float4); ALLOC(contactN, A.NC, float4);
    ALLOC(collideRes, A.NC, unsigned char);
    if ((d.reserved & 4) && A.NC <= 65535) ALLOC(collideItems, (size_t)2 * A.NC, uint32_t);  // mixed shape types: type-binned narrow phase
    ALLOC(movedList, A.NP, int);
    ALLOC(islandBodies, A.NB + A.NC, int); ALLOC(islandBodyEnd, A.NB, int);
    ALLOC(islandContacts, A.NC, int); ALLOC(islandContactEnd, A.NB, int);
    if (b->tileSolver) {
        ALLOC(velRec, (size_t)A.NC * kVelRec, float4); ALLOC(posRec, (size_t)A.NC * kPosRec, float4);
        ALLOC(impRec, A.NC, float4);
        ALLOC(solveVel, (size_t)A.NB * 3, float); ALLOC(solvePos, (size_t)A.NB * 3, float);
        ALLOC(solveCount, 1, int); ALLOC(solveFlags, 1, int); ALLOC(islandSolved, (A.NB + 31) / 32, int);
    } else {
        ALLOC(vc, (size_t)A.NC * kVcStride, float4);
    }
    ALLOC(toiMinAlpha, 1, float); ALLOC(contactToi, A.NC, float);
    ALLOC(toiCand, A.NC, int); ALLOC(toiCandCount, 1, int);
    ALLOC(bodyProxyStart, A.NB + 1, int); ALLOC(bodyProxyList, A.NP, int); ALLOC(toiBodies, kToiBodyCap, int); ALLOC(toiContacts, kMaxToiContacts, int);
    ALLOC(toiVc, (size_t)kMaxToiContacts * kVcStride, float4);
    ALLOC(eventA, A.NE, int4); ALLOC(eventB, A.NE, float4);
    ALLOC(jointCount, 1, int); ALLOC(jointSeq, 1, int); ALLOC(jointI, A.NJ, int4); ALLOC(jointF0, A.NJ, float4); ALLOC(jointF1, A.NJ, float4);
    ALLOC(jointF2, A.NJ, float4); ALLOC(jointImp, A.NJ, float4); ALLOC(jointTmp, (size_t)A.NJ * kJointTmp, float4);
    ALLOC(islandJoints, A.NJ, int); ALLOC(islandJointEnd, A.NB, int);

    b->fields.assign(B2GPU_F_COUNT, Field{nullptr, 0, 0});
    auto F = [&](int id, void** p, int cap, int comps) { b->fields[id] = Field{p, cap, comps}; };
    F(B2GPU_F_GRAVITY, (void**)&A.gravity, 1, 2);
    F(B2GPU_F_WORLD_FLAGS, (void**)&A.worldFlags, 1, 1);
    F(B2GPU_F_INV_DT0, (void**)&A.invDt0, 1, 1);
    F(B2GPU_F_BODY_COUNT, (void**)&A.bodyCount, 1, 1);
    F(B2GPU_F_PROXY_COUNT, (void**)&A.proxyCount, 1, 1);
    F(B2GPU_F_CONTACT_COUNT, (void**)&A.contactCount, 1, 1);
    F(B2GPU_F_STATUS, (void**)&A.status, 1, 1);
    F(B2GPU_F_EVENT_COUNT, (void**)&A.eventCount, 1, 1);
    F(B2GPU_F_ISLAND_COUNT, (void**)&A.islandCount, 1, 1);
    F(B2GPU_F_STEP_COUNT, (void**)&A.stepCount, 1, 1);
    F(B2GPU_F_BODY_XF, (void**)&A.bodyXf, A.NB, 4);
    F(B2GPU_F_BODY_SWEEP, (void**)&A.bodySweep, A.NB, 4);
    F(B2GPU_F_BODY_SWEEP0, (void**)&A.bodySweep0, A.NB, 4);
    F(B2GPU_F_BODY_VEL, (void**)&A.bodyVel, A.NB, 4);
    F(B2GPU_F_BODY_FORCE, (void**)&A.bodyForce, A.NB, 4);
    F(B2GPU_F_BODY_MASS, (void**)&A.bodyMass, A.NB, 4);
    F(B2GPU_F_BODY_DAMP, (void**)&A.bodyDamp, A.NB, 4);
    F(B2GPU_F_BODY_FLAGS, (void**)&A.bodyFlags, A.NB, 1);
    F(B2GPU_F_FIXTURE_INFO, (void**)&A.fixtureInfo, A.NF, 1);
    F(B2GPU_F_FIXTURE_SHAPE, (void**)&A.fixtureShape, A.NF, 1);
    F(B2GPU_F_FIXTURE_MAT, (void**)&A.fixtureMat, A.NF, 2);
    F(B2GPU_F_FIXTURE_FILTER, (void**)&A.fixtureFilter, A.NF, 2);
    F(B2GPU_F_PROXY_FAT, (void**)&A.proxyFat, A.NP, 4);
    F(B2GPU_F_PROXY_INFO, (void**)&A.proxyInfo, A.NP, 1);
    F(B2GPU_F_PROXY_MOVED, (void**)&A.proxyMoved, A.NP, 1);
    F(B2GPU_F_SHAPE_WORDS, (void**)&A.shapeWords, A.NS, 1);
    F(B2GPU_F_CHAIN_VERTS, (void**)&A.chainVerts, A.NV, 2);
    F(B2GPU_F_CONTACT_A, (void**)&A.contactA, A.NC, 4);
    F(B2GPU_F_CONTACT_K, (void**)&A.contactK, A.NC, 4);
    F(B2GPU_F_CONTACT_S, (void**)&A.contactS, A.NC, 2);
    F(B2GPU_F_CONTACT_P0, (void**)&A.contactP0, A.NC, 4);
    F(B2GPU_F_CONTACT_P1, (void**)&A.contactP1, A.NC, 4);
    F(B2GPU_F_CONTACT_N, (void**)&A.contactN, A.NC, 4);
    F(B2GPU_F_EVENT_A, (void**)&A.eventA, A.NE, 4);
    F(B2GPU_F_EVENT_B, (void**)&A.eventB, A.NE, 4);
    F(B2GPU_F_ISLAND_CONTACTS, (void**)&A.islandContacts, A.NC, 1);
    F(B2GPU_F_ISLAND_CONTACT_END, (void**)&A.islandContactEnd, A.NB, 1);
    F(B2GPU_F_JOINT_COUNT, (void**)&A.jointCount, 1, 1);
    F(B2GPU_F_JOINT_I, (void**)&A.jointI, A.NJ, 4);
    F(B2GPU_F_JOINT_F0, (void**)&A.jointF0, A.NJ, 4);
    F(B2GPU_F_JOINT_F1, (void**)&A.jointF1, A.NJ, 4);
    F(B2GPU_F_JOINT_F2, (void**)&A.jointF2, A.NJ, 4);
    F(B2GPU_F_JOINT_IMPULSE, (void**)&A.jointImp, A.NJ, 4);
    F(B2GPU_F_FIXTURE_COUNT, (void**)&A.fixtureCount, 1, 1);
    F(B2GPU_F_SHAPE_WORD_COUNT, (void**)&A.shapeWordCount, 1, 1);
    F(B2GPU_F_CHAIN_VERT_COUNT, (void**)&A.chainVertCount, 1, 1);
    F(B2GPU_F_PROXY_FREE_COUNT, (void**)&A.proxyFreeCount, 1, 1);
    F(B2GPU_F_PROXY_FREE, (void**)&A.proxyFree, A.NP, 1);
    F(B2GPU_F_NEW_FIXTURE, (void**)&A.newFixture, 1, 1);
    F(B2GPU_F_JOINT_SEQ, (void**)&A.jointSeq, 1, 1);
    F(B2GPU_F_FIXTURE_MASS, (void**)&A.fixtureMass, A.NF, 4);
    return B2GPU_OK;
}

int transferField(b2gpu_batch* b, int field, void* host, int firstWorld, int worldCount, bool toArena) {
    if (!b || field < 0 || field >= B2GPU_F_COUNT || !host) return fail(B2GPU_ERR_INVALID, "bad field transfer arguments");
    if (firstWorld < 0 || worldCount < 0 || firstWorld + worldCount > b->desc.numWorlds)
        return fail(B2GPU_ERR_INVALID, "world range [%d, %d) outside the batch", firstWorld, firstWorld + worldCount);
    if (worldCount == 0) return B2GPU_OK;
    CUDA_TRY(cudaSetDevice(b->desc.device));
    const Field& f = b->fields[field];
    size_t words = (size_t)worldCount * f.capacity * f.components;
    size_t bytes = words * 4;
    uint32_t* staging = nullptr;
    CUDA_TRY(cudaMalloc((void**)&staging, bytes));
    int threads = 256;
    unsigned blocks = (unsigned)((words + threads - 1) / threads);
    cudaError_t e = cudaSuccess;
    if (toArena) {
        e = cudaMemcpyAsync(staging, host, bytes, cudaMemcpyHostToDevice, b->stream);
        if (e == cudaSuccess) {
            k_interleave<<<blocks, threads, 0, b->stream>>>((uint32_t*)*f.dev, staging, f.capacity, f.components, firstWorld,
                                                            worldCount, 1);
            e = cudaGetLastError();
        }
    } else {
        k_interleave<<<blocks, threads, 0, b->stream>>>((uint32_t*)*f.dev, staging, f.capacity, f.components, firstWorld,
                                                        worldCount, 0);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, staging, bytes, cudaMemcpyDeviceToHost, b->stream);
    }
    b->launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    cudaFree(staging);
    if (e != cudaSuccess) return fail(B2GPU_ERR_CUDA, "field transfer failed: %s", cudaGetErrorString(e));
    return B2GPU_OK;
}

// Shared-memory configuration of the step kernels for this batch's per-world capacities.
int configureKernels(b2gpu_batch* b) {
    const b2gpu_batch_desc* desc = &b->desc;
    int maxSmem = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, desc->device));
    // Solver path: the tile path keeps the b2Velocity / b2Position table of 32 worlds in shared memory.
    b->solveSmem = SolveSmemBytes(desc->maxBodies);
    b->tileSolver = b->solveSmem <= (size_t)maxSmem && desc->maxBodies < 32768 && (desc->reserved & 1) == 0;
    b->collideSmemOk = (size_t)desc->maxContacts * kTile <= (size_t)maxSmem - 2048;
    if (b->collideSmemOk)
        CUDA_TRY(cudaFuncSetAttribute(k_collide<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, desc->maxContacts * kTile));
    if (b->tileSolver) {
        CUDA_TRY(cudaFuncSetAttribute(k_solve_velocity<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->solveSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_solve_velocity<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->solveSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_solve_position<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->solveSmem));
        CUDA_TRY(cudaFuncSetAttribute(k_solve_position<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->solveSmem));
    }
    // k_islands: as many worlds per CTA as keep two or more CTAs resident per SM (the DFS is latency bound)
    const size_t perWorld = IslandSmemPerWorld(desc->maxBodies, desc->maxContacts, desc->maxJoints > 0 ? desc->maxJoints : 0);
    int L = 32;
    size_t islandBudget = (size_t)100 * 1024;
    if (const char* env = getenv("B2GPU_ISLAND_KB")) {  // experiments: shared memory per k_islands CTA
        const int kb = atoi(env);
        if (kb >= 1 && kb <= 200) islandBudget = (size_t)kb * 1024;
    }
    while (L > 1 && perWorld * L > islandBudget) L >>= 1;
    if (perWorld * L > (size_t)maxSmem - 1024)
        return fail(B2GPU_ERR_CAPACITY, "a world with %d bodies / %d contacts exceeds the island kernel's shared memory",
                    desc->maxBodies, desc->maxContacts);
    b->islandLanes = L;
    b->islandSmem = perWorld * L;
    {  // persistent grid: as many CTAs as are resident at once
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, desc->device);
        int perSm = (int)((size_t)(maxSmem + 1024) / (b->islandSmem + 1024));
        if (perSm < 1) perSm = 1;
        if (perSm > 8) perSm = 8;
        b->islandGrid = sms * perSm;
    }
    CUDA_TRY(cudaFuncSetAttribute(k_islands, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->islandSmem));
    return B2GPU_OK;
}

// Kernel names in launch order of one step (b2gpu_batch_profile_kernels reports times in this order).
const char* const kTileKernels[] = {"k_collide", "k_islands", "k_solve_prepare", "k_solve_velocity", "k_solve_integrate",
                                    "k_solve_position", "k_solve_finalize", "k_sync_pairs", "k_toi", "k_toi_events"};
const char* const kGenericKernels[] = {"k_collide", "k_islands", "k_solve", "k_sync_pairs", "k_toi", "k_toi_events"};
constexpr int kMaxStepKernels = 10;

// k_toi only runs when some world has continuous physics on (b2World::Step: `if (m_continuousPhysics && dt > 0)`)
int stepKernelCount(const b2gpu_batch* b) { return (b->tileSolver ? 8 : 4) + (b->anyContinuous ? 2 : 0); }

// The kernel chain of one step for tiles [tile0, tile0 + tileCount) on stream st, in three parts:
//   part 0  k_collide, k_islands, k_solve_prepare            parallel kernels, little shared memory
//   part 1  k_solve_velocity, k_solve_integrate, k_solve_position   the sequential kernels: they own the SMs' shared memory
//   part 2  k_solve_finalize, k_sync_pairs, k_toi, k_toi_events
// (the generic solver path has k_solve as part 1).  `m` is the running index into marks.
void enqueuePart(b2gpu_batch* b, int part, float dt, int velIters, int posIters, cudaEvent_t* marks, int* m, cudaStream_t st,
                 int tile0, int tileCount) {
    Arena A = b->A;
    A.tile0 = tile0;
    const int threads = b->blockSize;
    const unsigned tiles = (unsigned)tileCount;
    const unsigned blocks = (unsigned)(((size_t)tileCount * kTile + threads - 1) / threads);
    auto mark = [&]() { if (marks) cudaEventRecord(marks[(*m)++], st); };
    if (part == 0) {
        mark();
        // result bytes in shared memory when the batch steps as one group, in global memory when groups co-reside
        if (b->groups <= 1 && b->collideSmemOk) k_collide<true><<<tiles, kCollideWarps * 32, (size_t)A.NC * kTile, st>>>(A);
        else k_collide<false><<<tiles, kCollideWarps * 32, 0, st>>>(A);
        mark();
        if (dt > 0.0f) {
            k_islands<<<tiles, kIslandThreads, b->islandSmem, st>>>(A, b->islandLanes, tileCount);
            mark();
        }
    } else if (dt <= 0.0f) {
        return;
    } else if (part == 3) {  // the last kernel of part 0, on its own so that late forces can be slipped in front of it
        if (b->tileSolver) {
            k_solve_prepare<<<tiles, kPrepWarps * 32, 0, st>>>(A, dt);
            mark();
        }
    } else if (part == 1) {
        if (b->tileSolver) {
            if (A.hasJoints) k_solve_velocity<true><<<tiles, 128, b->solveSmem, st>>>(A, velIters, dt);
            else k_solve_velocity<false><<<tiles, 128, b->solveSmem, st>>>(A, velIters, dt);
            mark();
            k_solve_integrate<<<tiles, kPrepWarps * 32, 0, st>>>(A, dt);
            mark();
            if (A.hasJoints) k_solve_position<true><<<tiles, 128, b->solveSmem, st>>>(A, posIters);
            else k_solve_position<false><<<tiles, 128, b->solveSmem, st>>>(A, posIters);
            mark();
        } else {
            k_solve<<<blocks, threads, 0, st>>>(A, dt, velIters, posIters);
            mark();
        }
    } else if (part == 4) {  // the first kernel of part 2, on its own: the poses are final behind it unless k_toi follows
        if (b->tileSolver) {
            k_solve_finalize<<<tiles, kPrepWarps * 32, 0, st>>>(A, dt);
            mark();
        }
    } else {
        k_sync_pairs<<<tiles, kPairWarps * 32, 0, st>>>(A, dt);
        mark();
        if (b->anyContinuous) {  // if (m_continuousPhysics && step.dt > 0) SolveTOI(step)
            k_toi<<<tiles, kToiWarps * 32, 0, st>>>(A, dt, velIters);
            mark();
            k_toi_events<<<tiles, 32, 0, st>>>(A, dt, velIters);
            mark();
        }
    }
}

// `count` b2World::Step calls for the whole batch (Box2D/Dynamics/b2World.cpp:888-954).
// marks (optional): per chain stepKernelCount + 1 events, one before the first kernel and one after each.
//
// World groups (b->groups > 1): every group of tiles has its own stream and runs the same chain.  The sequential
// kernels of a tile solver step hold nearly all shared memory of the SMs they run on (2 x 111 KB) but issue from two
// warps per SM, while the parallel kernels need registers and issue slots but no shared memory: the groups are
// software-pipelined so that one group's part 1 runs while the other groups are in parts 0 / 2 — group g's part 1 of
// a step waits for group g-1's part 1 of the same step (group 0: for the last group's part 1 of the previous step).
//
// Per-step host I/O (io != nullptr, b2gpu_batch_step_io): every group copies ITS worlds' forces in before its chain and
// its worlds' transforms out after it, on its own stream, so one group's PCIe transfers run under the other groups'
// kernels instead of in front of and behind the whole batch's step.
struct StepIo {
    const float* forces;   // host, [world][ioForceCount][3] or null
    float* transforms;     // host, [world][ioPoseCount][3 or 4] or null
    float* devForces;      // device staging, same shapes
    float* devTransforms;
    int groups;            // world groups for this call (>= b->groups)
};

// Streams and events of groups [0, G) (created on first use: b2gpu_batch_step_io may use more groups than the step).
int ensureGroupStreams(b2gpu_batch* b, int G) {
    for (int g = 0; g < G && g < b2gpu_batch::kMaxGroups; ++g) {
        if (g > 0 && !b->groupStream[g]) CUDA_TRY(cudaStreamCreateWithFlags(&b->groupStream[g], cudaStreamNonBlocking));
        if (g > 0 && !b->evJoin[g]) CUDA_TRY(cudaEventCreateWithFlags(&b->evJoin[g], cudaEventDisableTiming));
        if (!b->evSeq[g]) CUDA_TRY(cudaEventCreateWithFlags(&b->evSeq[g], cudaEventDisableTiming));
        if (!b->copyStream[g]) CUDA_TRY(cudaStreamCreateWithFlags(&b->copyStream[g], cudaStreamNonBlocking));
        if (!b->evCopy[g]) CUDA_TRY(cudaEventCreateWithFlags(&b->evCopy[g], cudaEventDisableTiming));
    }
    return B2GPU_OK;
}

int enqueueSteps(b2gpu_batch* b, float dt, int velIters, int posIters, cudaEvent_t* marks, int count, bool traceGroups = false,
                 const StepIo* io = nullptr) {
    cudaStream_t st = b->stream;
    if (b->needInitialPairs) {  // if (m_flags & e_newFixture) FindNewContacts()
        const int threads = b->blockSize;
        const unsigned blocks = (unsigned)(((size_t)b->A.numTiles * kTile + threads - 1) / threads);
        k_find_new_contacts<<<blocks, threads, 0, st>>>(b->A);
        b->needInitialPairs = false;
        b->launches++;
    }
    if (b->needProxyIndex) {  // body -> proxies index (derived state: after the worlds were built or restored)
        k_index_proxies<<<(unsigned)((b->A.numTiles * kTile + 127) / 128), 128, 0, st>>>(b->A);
        b->needProxyIndex = false;
        b->launches++;
    }
    const int perStep = dt > 0.0f ? stepKernelCount(b) : 1;
    const int nkTrace = stepKernelCount(b) + 1;  // traceGroups: marks is [group][step][nk + 1]
    int G = ((marks && !traceGroups) || b->groups < 1) ? 1 : b->groups;
    if (io) G = io->groups;
    if (G > b->A.numTiles) G = b->A.numTiles > 0 ? b->A.numTiles : 1;
    const int per = (b->A.numTiles + G - 1) / G;
    // experiment (B2GPU_GROUP0_TILES): with two groups, group 0 takes this many tiles (e.g. one full wave of resident tiles)
    int first = per;
    if (G == 2 && !io) {
        if (const char* env = getenv("B2GPU_GROUP0_TILES")) {
            const int v = atoi(env);
            if (v > 0 && v < b->A.numTiles) first = v;
        }
    }
    if (G > 1 || io) {  // fork: every group stream waits for what is already queued on the batch's stream
        cudaEventRecord(b->evFork, st);
        for (int g = 1; g < G; ++g) cudaStreamWaitEvent(b->groupStream[g], b->evFork, 0);
    }
    const bool pipeline = G > 1 && b->tileSolver && dt > 0.0f && b->pipelineGroups;
    for (int i = 0; i < count; ++i) {
        for (int g = 0; g < G; ++g) {
            const int t0 = (G == 2 && g == 1) ? first : g * per;
            const int n = (G == 2 && !io) ? (g == 0 ? first : b->A.numTiles - first)
                                          : (b->A.numTiles - t0 < per ? b->A.numTiles - t0 : per);
            if (n <= 0) break;
            cudaStream_t gs = g == 0 ? st : b->groupStream[g];
            cudaEvent_t* mk = marks ? marks + (traceGroups ? ((size_t)g * count + i) * nkTrace : 0) : nullptr;
            int m = 0;
            // this group's worlds and their slice of the host buffers
            const int w0 = t0 * kTile;
            const int w1 = (t0 + n) * kTile < b->desc.numWorlds ? (t0 + n) * kTile : b->desc.numWorlds;
            const int nF = b->ioForceCount, nP = b->ioPoseCount, poseFloats = b->ioPoseFormat == B2GPU_POSE_XYA ? 3 : 4;
            Arena Ag = b->A;
            Ag.tile0 = t0;
            const bool haveForces = io && io->forces && w1 > w0 && nF > 0;
            // Forces only matter from the velocity integration on (k_solve_prepare / k_solve) — unless one wakes a sleeping
            // body, which k_collide and k_islands would have to see.  No sleepers possible: the copy runs on its own stream
            // under k_collide / k_islands and the forces are applied behind them.
            const bool lateForces = haveForces && !b->maySleep && dt > 0.0f && b->copyStream[g] != nullptr;
            const size_t fOff = (size_t)w0 * nF * 3, fLen = (size_t)(w1 - w0) * nF * 3;
            const size_t fItems = (size_t)n * nF * kTile;
            if (haveForces && !lateForces) {
                cudaMemcpyAsync(io->devForces + fOff, io->forces + fOff, fLen * sizeof(float), cudaMemcpyHostToDevice, gs);
                k_apply_forces<<<(unsigned)((fItems + 255) / 256), 256, 0, gs>>>(Ag, io->devForces, n, b->ioForceSel, nF);
                b->launches++;
            }
            if (lateForces) {
                cudaStreamWaitEvent(b->copyStream[g], b->evFork, 0);  // not before what the batch's stream held at the fork
                cudaMemcpyAsync(io->devForces + fOff, io->forces + fOff, fLen * sizeof(float), cudaMemcpyHostToDevice, b->copyStream[g]);
                cudaEventRecord(b->evCopy[g], b->copyStream[g]);
            }
            enqueuePart(b, 0, dt, velIters, posIters, mk, &m, gs, t0, n);
            if (lateForces) {
                cudaStreamWaitEvent(gs, b->evCopy[g], 0);
                k_apply_forces<<<(unsigned)((fItems + 255) / 256), 256, 0, gs>>>(Ag, io->devForces, n, b->ioForceSel, nF);
                b->launches++;
            }
            enqueuePart(b, 3, dt, velIters, posIters, mk, &m, gs, t0, n);
            if (pipeline) {
                const int prev = g == 0 ? G - 1 : g - 1;
                if (g > 0 || i > 0) cudaStreamWaitEvent(gs, b->evSeq[prev], 0);
            }
            enqueuePart(b, 1, dt, velIters, posIters, mk, &m, gs, t0, n);
            if (pipeline) cudaEventRecord(b->evSeq[g], gs);
            enqueuePart(b, 4, dt, velIters, posIters, mk, &m, gs, t0, n);
            enqueuePart(b, 2, dt, velIters, posIters, mk, &m, gs, t0, n);
            // (reading the poses out on the side stream right behind k_solve_finalize, under k_sync_pairs, was measured: no
            // gain with 4 groups, a loss with more)
            if (io && io->transforms && w1 > w0 && nP > 0) {
                const size_t off = (size_t)w0 * nP * poseFloats, len = (size_t)(w1 - w0) * nP * poseFloats;
                const size_t items = (size_t)n * nP * kTile;
                k_read_transforms<<<(unsigned)((items + 255) / 256), 256, 0, gs>>>(Ag, io->devTransforms, n, b->ioPoseSel, nP,
                                                                                 b->ioPoseFormat);
                b->launches++;
                cudaMemcpyAsync(io->transforms + off, io->devTransforms + off, len * sizeof(float), cudaMemcpyDeviceToHost, gs);
            }
        }
    }
    for (int g = 1; g < G; ++g) {  // join
        cudaEventRecord(b->evJoin[g], b->groupStream[g]);
        cudaStreamWaitEvent(st, b->evJoin[g], 0);
    }
    b->launches += (int64_t)perStep * G * count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(B2GPU_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return B2GPU_OK;
}
// Captured graphs bake in the kernel chain (k_toi / k_toi_events only when some world steps continuously), the world
// groups and the step_io staging pointers: whatever changes one of those drops every graph, they are re-captured on use.
void invalidateGraphs(b2gpu_batch* b) {
    if (b->graphExecIo) { cudaGraphExecDestroy(b->graphExecIo); b->graphExecIo = nullptr; }
    if (b->graphExec) { cudaGraphExecDestroy(b->graphExec); b->graphExec = nullptr; }
    if (b->graphExecChunk) { cudaGraphExecDestroy(b->graphExecChunk); b->graphExecChunk = nullptr; }
}

int enqueueStep(b2gpu_batch* b, float dt, int velIters, int posIters, cudaEvent_t* marks) {
    return enqueueSteps(b, dt, velIters, posIters, marks, 1);
}

}  // namespace

extern "C" {

const char* b2gpu_last_error(void) { return g_error.c_str(); }

// lets the host layer (b2g_host.cpp) report through the same thread-local message
int b2gpu_host_set_error(int code, const char* msg) { return fail(code, "%s", msg); }

// How many world groups step concurrently (measured with tools/groups_sweep2.sh, pyramid-20, ms per step for 1 / 2 / 4
// groups): 8,192 worlds 3.46 / 3.47 / 3.56; 16,384 worlds 5.75 / 5.31 / 5.32; 32,768 worlds 10.79 / 9.55 / 9.14.
// Groups pay off once the sequential kernels need more than one wave of resident tiles (2 per SM).
static int defaultGroups(const b2gpu_batch* b) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->desc.device);
    const int wave = 2 * sms;
    if (!b->tileSolver || b->A.numTiles <= wave) return 1;
    return b->A.numTiles >= 3 * wave ? 4 : 2;
}

int b2gpu_batch_create(const b2gpu_batch_desc* desc, b2gpu_batch** out) {
    if (!desc || !out) return fail(B2GPU_ERR_INVALID, "null argument");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B2GPU_ERR_NO_DEVICE, "no CUDA device available (%s); box2dpp_b200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (desc->device < 0 || desc->device >= count) return fail(B2GPU_ERR_INVALID, "device %d out of range", desc->device);
    if (desc->numWorlds <= 0 || desc->maxBodies <= 0 || desc->maxFixtures <= 0 || desc->maxProxies <= 0 ||
        desc->maxContacts <= 0 || desc->maxShapeWords <= 0)
        return fail(B2GPU_ERR_INVALID, "capacities must be positive");
    if (desc->maxBodies > 65535 || desc->maxFixtures > 65535 || desc->maxProxies > 65535)
        return fail(B2GPU_ERR_CAPACITY, "bodies, fixtures and proxies per world are limited to 65535");
    if (desc->maxJoints < 0 || desc->maxEvents < 0 || desc->maxChainVerts < 0)
        return fail(B2GPU_ERR_INVALID, "maxJoints, maxEvents and maxChainVerts must not be negative");
    {   // the arenas index with size_t, but the per-world capacities multiply: refuse what cannot be a real request
        const double perWorld = 200.0 * desc->maxBodies + 48.0 * desc->maxFixtures + 40.0 * desc->maxProxies +
                                600.0 * desc->maxContacts + 4.0 * desc->maxShapeWords + 8.0 * desc->maxChainVerts +
                                32.0 * desc->maxEvents + 200.0 * desc->maxJoints;
        if (perWorld * (double)desc->numWorlds > 1.0e12)
            return fail(B2GPU_ERR_CAPACITY, "numWorlds x capacities asks for more than 1 TB of device memory");
    }
    if (desc->blockSize != 0 && (desc->blockSize % 32 != 0 || desc->blockSize > 128))
        return fail(B2GPU_ERR_INVALID, "blockSize must be 32, 64, 96 or 128");
    CUDA_TRY(cudaSetDevice(desc->device));
    b2gpu_batch* b = new b2gpu_batch;
    b->desc = *desc;
    b->blockSize = desc->blockSize ? desc->blockSize : 32;
    memset(&b->A, 0, sizeof(b->A));
    {
        int rc = configureKernels(b);
        if (rc != B2GPU_OK) { delete b; return rc; }
    }
    e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete b; return fail(B2GPU_ERR_CUDA, "stream create failed: %s", cudaGetErrorString(e)); }
    int rc = allocateArena(b);
    b->ioForceCount = b->ioPoseCount = b->A.NB;  // dense I/O covers every body slot until b2gpu_batch_io_select says otherwise
    if (rc == B2GPU_OK) {
        e = cudaEventCreate(&b->evStart);
        if (e == cudaSuccess) e = cudaEventCreate(&b->evStop);
        // World groups (see b2gpu_batch::groups): B2GPU_GROUPS overrides the default for experiments.
        b->groups = defaultGroups(b);
        if (const char* env = getenv("B2GPU_GROUPS")) {
            const int g = atoi(env);
            if (g >= 1 && g <= b2gpu_batch::kMaxGroups) b->groups = g;
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->evFork, cudaEventDisableTiming);
        if (const char* env = getenv("B2GPU_PIPELINE")) b->pipelineGroups = atoi(env) != 0;
        for (int g = 0; g < b->groups && e == cudaSuccess; ++g) {
            if (g > 0) e = cudaStreamCreateWithFlags(&b->groupStream[g], cudaStreamNonBlocking);
            if (e == cudaSuccess && g > 0) e = cudaEventCreateWithFlags(&b->evJoin[g], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->evSeq[g], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
        if (e != cudaSuccess) rc = fail(B2GPU_ERR_CUDA, "batch setup failed: %s", cudaGetErrorString(e));
    }
    if (rc != B2GPU_OK) { b2gpu_batch_destroy(b); return rc; }
    *out = b;
    return B2GPU_OK;
}

int b2gpu_batch_destroy(b2gpu_batch* b) {
    if (!b) return B2GPU_OK;
    cudaSetDevice(b->desc.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (void* p : b->allocations) cudaFree(p);
    if (b->ioDev) cudaFree(b->ioDev);
    if (b->ioForceSel) cudaFree(b->ioForceSel);
    if (b->ioPoseSel) cudaFree(b->ioPoseSel);
    if (b->cmdDev) cudaFree(b->cmdDev);
    for (int g = 0; g < b2gpu_batch::kMaxGroups; ++g) {
        if (b->copyStream[g]) cudaStreamDestroy(b->copyStream[g]);
        if (b->evCopy[g]) cudaEventDestroy(b->evCopy[g]);
    }
    if (b->graphExec) cudaGraphExecDestroy(b->graphExec);
    if (b->graphExecChunk) cudaGraphExecDestroy(b->graphExecChunk);
    if (b->graphExecIo) cudaGraphExecDestroy(b->graphExecIo);
    for (int g = 1; g < b2gpu_batch::kMaxGroups; ++g) {
        if (b->groupStream[g]) { cudaStreamSynchronize(b->groupStream[g]); cudaStreamDestroy(b->groupStream[g]); }
        if (b->evJoin[g]) cudaEventDestroy(b->evJoin[g]);
    }
    for (int g = 0; g < b2gpu_batch::kMaxGroups; ++g)
        if (b->evSeq[g]) cudaEventDestroy(b->evSeq[g]);
    if (b->evFork) cudaEventDestroy(b->evFork);
    if (b->evStart) cudaEventDestroy(b->evStart);
    if (b->evStop) cudaEventDestroy(b->evStop);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return B2GPU_OK;
}

int b2gpu_batch_upload_field(b2gpu_batch* b, int32_t field, const void* host, int32_t firstWorld, int32_t worldCount) {
    int rc = transferField(b, field, const_cast<void*>(host), firstWorld, worldCount, true);
    // a freshly uploaded move buffer == b2World::e_newFixture: the next Step starts with FindNewContacts
    if (rc == B2GPU_OK && field == B2GPU_F_PROXY_MOVED) b->needInitialPairs = true;
    if (rc == B2GPU_OK && (field == B2GPU_F_PROXY_INFO || field == B2GPU_F_FIXTURE_INFO)) b->needProxyIndex = true;
    if (rc == B2GPU_OK && field == B2GPU_F_WORLD_FLAGS && !b->anyContinuous) {
        const int32_t* flags = static_cast<const int32_t*>(host);
        for (int32_t i = 0; i < worldCount; ++i)
            if (flags[i] & B2G_WORLD_CONTINUOUS) { b->anyContinuous = true; break; }
        if (b->anyContinuous) {  // the captured steps have no k_toi: capture again
            invalidateGraphs(b);
        }
    }
    // can a dynamic body be asleep from now on?  (b2gpu_batch::maySleep: the graph of step_io bakes the answer in)
    if (rc == B2GPU_OK && !b->maySleep && (field == B2GPU_F_WORLD_FLAGS || field == B2GPU_F_BODY_FLAGS)) {
        const int32_t* v = static_cast<const int32_t*>(host);
        const size_t nWords = (size_t)worldCount * b->fields[field].capacity;
        for (size_t i = 0; i < nWords && !b->maySleep; ++i) {
            if (field == B2GPU_F_WORLD_FLAGS) b->maySleep = (v[i] & B2G_WORLD_ALLOW_SLEEP) != 0;
            else b->maySleep = (v[i] & kBodyTypeMask) == B2G_DYNAMIC_BODY && (v[i] & kBodyAwake) == 0 && (v[i] & kBodyDestroyed) == 0;
        }
        if (b->maySleep) invalidateGraphs(b);
    }
    // any host edit may change what the islands depend on: rebuild them on the next step
    if (rc == B2GPU_OK) cudaMemsetAsync(b->A.topoDirty, 1, (size_t)b->A.numTiles * kTile * sizeof(int), b->stream);
    return rc;
}
int b2gpu_batch_download_field(b2gpu_batch* b, int32_t field, void* host, int32_t firstWorld, int32_t worldCount) {
    return transferField(b, field, host, firstWorld, worldCount, false);
}
int b2gpu_batch_field_shape(b2gpu_batch* b, int32_t field, int32_t* capacity, int32_t* components) {
    if (!b || field < 0 || field >= B2GPU_F_COUNT) return fail(B2GPU_ERR_INVALID, "bad field id");
    if (capacity) *capacity = b->fields[field].capacity;
    if (components) *components = b->fields[field].components;
    return B2GPU_OK;
}

int b2gpu_batch_step(b2gpu_batch* b, float dt, int32_t velIters, int32_t posIters) {
    if (!b) return fail(B2GPU_ERR_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    return enqueueStep(b, dt, velIters, posIters, nullptr);
}

// Captures `count` steps into a graph (all group streams join the capture through the fork / join events).
static int captureSteps(b2gpu_batch* b, float dt, int velIters, int posIters, int count, cudaGraphExec_t* out) {
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
    int64_t before = b->launches;
    int rc = enqueueSteps(b, dt, velIters, posIters, nullptr, count);
    b->launches = before;
    cudaError_t e = cudaStreamEndCapture(b->stream, &graph);
    if (rc != B2GPU_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(B2GPU_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(out, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(B2GPU_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    return B2GPU_OK;
}

int b2gpu_batch_step_n(b2gpu_batch* b, float dt, int32_t velIters, int32_t posIters, int32_t steps) {
    if (!b || steps < 0) return fail(B2GPU_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    if (steps == 0) return B2GPU_OK;
    int done = 0;
    CUDA_TRY(cudaEventRecord(b->evStart, b->stream));
    if (b->needInitialPairs || b->needProxyIndex) {  // the first step also runs FindNewContacts: keep it out of the graph
        int rc = enqueueStep(b, dt, velIters, posIters, nullptr);
        if (rc != B2GPU_OK) return rc;
        done = 1;
    }
    if (done < steps) {
        if (b->graphDt != dt || b->graphVel != velIters || b->graphPos != posIters) {
            if (b->graphExec) { cudaGraphExecDestroy(b->graphExec); b->graphExec = nullptr; }
            if (b->graphExecChunk) { cudaGraphExecDestroy(b->graphExecChunk); b->graphExecChunk = nullptr; }
            b->graphDt = dt; b->graphVel = velIters; b->graphPos = posIters;
        }
        const int perStep = (dt > 0.0f ? stepKernelCount(b) : 1) * (b->groups > 1 ? b->groups : 1);
        // several groups: graphs of kGraphChunk steps, so that the groups only re-synchronise at chunk boundaries
        const int chunk = b->groups > 1 ? kGraphChunk : 1;
        if (chunk > 1 && steps - done >= chunk) {
            if (!b->graphExecChunk) {
                int rc = captureSteps(b, dt, velIters, posIters, chunk, &b->graphExecChunk);
                if (rc != B2GPU_OK) return rc;
            }
            for (; steps - done >= chunk; done += chunk) {
                CUDA_TRY(cudaGraphLaunch(b->graphExecChunk, b->stream));
                b->launches += (int64_t)perStep * chunk;
            }
        }
        if (done < steps) {
            if (!b->graphExec) {
                int rc = captureSteps(b, dt, velIters, posIters, 1, &b->graphExec);
                if (rc != B2GPU_OK) return rc;
            }
            for (; done < steps; ++done) {
                CUDA_TRY(cudaGraphLaunch(b->graphExec, b->stream));
                b->launches += perStep;
            }
        }
    }
    CUDA_TRY(cudaEventRecord(b->evStop, b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_sync(b2gpu_batch* b) {
    if (!b) return fail(B2GPU_ERR_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_last_step_ms(b2gpu_batch* b, float* ms) {
    if (!b || !ms) return fail(B2GPU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    CUDA_TRY(cudaEventSynchronize(b->evStop));
    CUDA_TRY(cudaEventElapsedTime(ms, b->evStart, b->evStop));
    return B2GPU_OK;
}

int b2gpu_batch_profile_kernels(b2gpu_batch* b, float dt, int32_t velIters, int32_t posIters, int32_t steps, float* ms,
                                int32_t capacity, int32_t* count) {
    if (!b || !ms || steps <= 0 || dt <= 0.0f) return fail(B2GPU_ERR_INVALID, "bad arguments");
    const int nk = stepKernelCount(b);
    if (capacity < nk) return fail(B2GPU_ERR_INVALID, "ms[] needs room for %d kernels", nk);
    CUDA_TRY(cudaSetDevice(b->desc.device));
    cudaEvent_t marks[kMaxStepKernels + 1];
    for (int i = 0; i <= nk; ++i) CUDA_TRY(cudaEventCreate(&marks[i]));
    for (int i = 0; i < nk; ++i) ms[i] = 0.0f;
    int rc = B2GPU_OK;
    for (int s = 0; s < steps && rc == B2GPU_OK; ++s) {
        rc = enqueueStep(b, dt, velIters, posIters, marks);
        if (rc != B2GPU_OK) break;
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) { rc = fail(B2GPU_ERR_CUDA, "step failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
        for (int i = 0; i < nk; ++i) {
            float t = 0.0f;
            cudaEventElapsedTime(&t, marks[i], marks[i + 1]);
            ms[i] += t;
        }
    }
    for (int i = 0; i <= nk; ++i) cudaEventDestroy(marks[i]);
    if (count) *count = nk;
    return rc;
}

// Diagnostic: `steps` steps with the batch's world groups, a CUDA event after every kernel of every group.
// ms[(group * steps + step) * (kernels + 1) + k] = completion time of kernel k - 1 of that step on that group's
// stream (k = 0: when the chain started), in ms since the first chain started.
int b2gpu_batch_trace_groups(b2gpu_batch* b, float dt, int32_t velIters, int32_t posIters, int32_t steps, float* ms,
                             int32_t capacity, int32_t* groupsOut, int32_t* kernelsOut) {
    if (!b || !ms || steps <= 0 || dt <= 0.0f) return fail(B2GPU_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    const int nk = stepKernelCount(b), G = b->groups < b->A.numTiles ? b->groups : 1;
    const size_t n = (size_t)G * steps * (nk + 1);
    if ((size_t)capacity < n) return fail(B2GPU_ERR_INVALID, "ms[] needs room for %zu values", n);
    std::vector<cudaEvent_t> ev(n);
    for (size_t i = 0; i < n; ++i) CUDA_TRY(cudaEventCreate(&ev[i]));
    int rc = enqueueSteps(b, dt, velIters, posIters, ev.data(), steps, true);
    if (rc == B2GPU_OK && cudaStreamSynchronize(b->stream) != cudaSuccess) rc = fail(B2GPU_ERR_CUDA, "trace failed");
    for (size_t i = 0; i < n && rc == B2GPU_OK; ++i) {
        ms[i] = 0.0f;
        cudaEventElapsedTime(&ms[i], ev[0], ev[i]);
    }
    for (size_t i = 0; i < n; ++i) cudaEventDestroy(ev[i]);
    if (groupsOut) *groupsOut = G;
    if (kernelsOut) *kernelsOut = nk;
    return rc;
}

const char* b2gpu_batch_kernel_name(b2gpu_batch* b, int32_t index) {
    if (!b || index < 0 || index >= stepKernelCount(b)) return "";
    return b->tileSolver ? kTileKernels[index] : kGenericKernels[index];
}

int b2gpu_batch_profile_steps(b2gpu_batch* b, float dt, int32_t velIters, int32_t posIters, int32_t steps, float* ms) {
    float k[kMaxStepKernels];
    int32_t nk = 0;
    int rc = b2gpu_batch_profile_kernels(b, dt, velIters, posIters, steps, k, kMaxStepKernels, &nk);
    if (rc != B2GPU_OK) return rc;
    ms[0] = k[0];
    ms[1] = k[1];
    ms[2] = 0.0f;
    for (int i = 2; i < nk - 1; ++i) ms[2] += k[i];
    ms[3] = k[nk - 1];
    return B2GPU_OK;
}

#ifdef B2G_TOI_PROFILE
int b2gpu_debug_toi_profile(unsigned long long* out16, int reset) {
    cudaMemcpyFromSymbol(out16, g_toiProf, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_toiProf, z, sizeof(z)); }
    return 0;
}
#endif

int64_t b2gpu_batch_launch_count(b2gpu_batch* b) { return b ? b->launches : 0; }

static int ensureIo(b2gpu_batch* b, size_t need = 0) {
    if (need == 0) need = (size_t)b->desc.numWorlds * b->A.NB * 4 * sizeof(float);
    if (b->ioBytes < need) {
        if (b->graphExecIo) { cudaGraphExecDestroy(b->graphExecIo); b->graphExecIo = nullptr; }  // it holds the old pointer
        if (b->ioDev) cudaFree(b->ioDev);
        b->ioDev = nullptr;
        CUDA_TRY(cudaMalloc((void**)&b->ioDev, need));
        b->ioBytes = need;
    }
    return B2GPU_OK;
}

int b2gpu_batch_io_select(b2gpu_batch* b, const int32_t* forceBodies, int32_t forceCount, const int32_t* poseBodies,
                          int32_t poseCount, int32_t poseFormat) {
    if (!b) return fail(B2GPU_ERR_INVALID, "null batch");
    if (poseFormat != B2GPU_POSE_XF && poseFormat != B2GPU_POSE_XYA) return fail(B2GPU_ERR_INVALID, "unknown pose format %d", poseFormat);
    if (forceCount < -1 || poseCount < -1 || (forceCount > 0 && !forceBodies) || (poseCount > 0 && !poseBodies))
        return fail(B2GPU_ERR_INVALID, "bad body selection");
    for (int i = 0; i < forceCount; ++i)
        if (forceBodies[i] < 0 || forceBodies[i] >= b->A.NB) return fail(B2GPU_ERR_INVALID, "force body %d out of range", forceBodies[i]);
    for (int i = 0; i < poseCount; ++i)
        if (poseBodies[i] < 0 || poseBodies[i] >= b->A.NB) return fail(B2GPU_ERR_INVALID, "pose body %d out of range", poseBodies[i]);
    CUDA_TRY(cudaSetDevice(b->desc.device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));  // nothing in flight uses the old lists
    auto set = [&](int** dev, int* count, const int32_t* host, int n) -> int {
        if (*dev) { cudaFree(*dev); *dev = nullptr; }
        if (n < 0) { *count = b->A.NB; return B2GPU_OK; }  // every body slot
        *count = n;
        if (n == 0) return B2GPU_OK;
        CUDA_TRY(cudaMalloc((void**)dev, (size_t)n * sizeof(int)));
        CUDA_TRY(cudaMemcpy(*dev, host, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
        return B2GPU_OK;
    };
    int rc = set(&b->ioForceSel, &b->ioForceCount, forceBodies, forceCount);
    if (rc != B2GPU_OK) return rc;
    rc = set(&b->ioPoseSel, &b->ioPoseCount, poseBodies, poseCount);
    if (rc != B2GPU_OK) return rc;
    b->ioPoseFormat = poseFormat;
    invalidateGraphs(b);  // the step_io graph holds the old lists and sizes
    return B2GPU_OK;
}

int b2gpu_batch_apply_forces(b2gpu_batch* b, const float* forces) {
    if (!b || !forces) return fail(B2GPU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    if (b->ioForceCount == 0) return B2GPU_OK;
    size_t bytes = (size_t)b->desc.numWorlds * b->ioForceCount * 3 * sizeof(float);
    int rc = ensureIo(b, bytes);
    if (rc != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(b->ioDev, forces, bytes, cudaMemcpyHostToDevice, b->stream));
    size_t total = (size_t)b->A.numTiles * b->ioForceCount * kTile;
    Arena A = b->A;
    A.tile0 = 0;
    k_apply_forces<<<(unsigned)((total + 255) / 256), 256, 0, b->stream>>>(A, b->ioDev, b->A.numTiles, b->ioForceSel, b->ioForceCount);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2GPU_OK;
}

int b2gpu_batch_read_transforms(b2gpu_batch* b, float* out) {
    if (!b || !out) return fail(B2GPU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    if (b->ioPoseCount == 0) return B2GPU_OK;
    const int poseFloats = b->ioPoseFormat == B2GPU_POSE_XYA ? 3 : 4;
    size_t bytes = (size_t)b->desc.numWorlds * b->ioPoseCount * poseFloats * sizeof(float);
    int rc = ensureIo(b, bytes);
    if (rc != B2GPU_OK) return rc;
    size_t total = (size_t)b->A.numTiles * b->ioPoseCount * kTile;
    Arena A = b->A;
    A.tile0 = 0;
    k_read_transforms<<<(unsigned)((total + 255) / 256), 256, 0, b->stream>>>(A, b->ioDev, b->A.numTiles, b->ioPoseSel, b->ioPoseCount,
                                                                             b->ioPoseFormat);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, b->ioDev, bytes, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

// One b2World::Step of every world with its per-step host I/O fused in (see StepIo).
int b2gpu_batch_step_io(b2gpu_batch* b, const float* forces, float dt, int32_t velIters, int32_t posIters, float* transforms) {
    if (!b) return fail(B2GPU_ERR_INVALID, "null batch");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    // device staging: forces first, transforms behind them (16-byte aligned for the float4 stores of B2GPU_POSE_XF)
    const size_t forceFloats = ((size_t)b->desc.numWorlds * b->ioForceCount * 3 + 3) & ~(size_t)3;
    const size_t poseFloats = (size_t)b->desc.numWorlds * b->ioPoseCount * 4;
    int rc = ensureIo(b, (forceFloats + poseFloats) * sizeof(float));
    if (rc != B2GPU_OK) return rc;
    StepIo io;
    io.forces = forces;
    io.transforms = transforms;
    io.devForces = b->ioDev;
    io.devTransforms = b->ioDev + forceFloats;
    // eight groups: the exposed transfers are the first group's copy in and the last group's copy out, an eighth of the
    // batch's each (16,384 pyramid worlds, ms per step for 2 / 4 / 8 groups: 6.81 / 6.46 / 6.28; separate calls 8.60)
    io.groups = 4;  // measured (tools/io_groups_r2b.sh, 16,384 pyramid worlds): 2 / 4 / 5 / 8 groups = 5.35 / 5.12 / 5.09 / 5.72 ms per step
    if (const char* env = getenv("B2GPU_IO_GROUPS")) {
        const int g = atoi(env);
        if (g >= 1 && g <= b2gpu_batch::kMaxGroups) io.groups = g;
    }
    rc = ensureGroupStreams(b, io.groups);
    if (rc != B2GPU_OK) return rc;
    // A caller's loop passes the same pinned buffers every step: the whole call (copies included) is then one graph
    // launch instead of ~100 stream operations.  Pageable buffers cannot be captured; they take the direct path.
    auto pinned = [](const void* p) {
        if (!p) return true;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    const bool graphable = !b->needInitialPairs && !b->needProxyIndex && pinned(forces) && pinned(transforms) && !getenv("B2GPU_IO_NO_GRAPH");
    if (graphable) {
        const bool same = b->graphExecIo && b->ioGraphForces == forces && b->ioGraphTransforms == transforms && b->ioGraphDt == dt &&
                          b->ioGraphVel == velIters && b->ioGraphPos == posIters && b->ioGraphGroups == io.groups;
        if (!same) {
            if (b->graphExecIo) { cudaGraphExecDestroy(b->graphExecIo); b->graphExecIo = nullptr; }
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(b->stream, cudaStreamCaptureModeThreadLocal));
            const int64_t before = b->launches;
            rc = enqueueSteps(b, dt, velIters, posIters, nullptr, 1, false, &io);
            b->ioGraphLaunches = b->launches - before;
            b->launches = before;
            cudaError_t e = cudaStreamEndCapture(b->stream, &graph);
            if (rc != B2GPU_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return fail(B2GPU_ERR_CUDA, "step_io graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&b->graphExecIo, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { b->graphExecIo = nullptr; return fail(B2GPU_ERR_CUDA, "step_io graph instantiate failed: %s", cudaGetErrorString(e)); }
            b->ioGraphForces = forces; b->ioGraphTransforms = transforms; b->ioGraphDt = dt;
            b->ioGraphVel = velIters; b->ioGraphPos = posIters; b->ioGraphGroups = io.groups;
        }
        CUDA_TRY(cudaGraphLaunch(b->graphExecIo, b->stream));
        b->launches += b->ioGraphLaunches;
    } else {
        rc = enqueueSteps(b, dt, velIters, posIters, nullptr, 1, false, &io);
        if (rc != B2GPU_OK) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(b->stream));  // joined: the caller holds this step's transforms
    return B2GPU_OK;
}

int b2gpu_batch_read_body_state(b2gpu_batch* b, float* out) {
    if (!b || !out) return fail(B2GPU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    const size_t bytes = (size_t)b->desc.numWorlds * b->A.NB * 8 * sizeof(float);
    int rc = ensureIo(b, bytes);
    if (rc != B2GPU_OK) return rc;
    size_t total = (size_t)b->A.numTiles * b->A.NB * kTile;
    k_read_body_state<<<(unsigned)((total + 255) / 256), 256, 0, b->stream>>>(b->A, (float4*)b->ioDev);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, b->ioDev, bytes, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

// ---- snapshot / restore ------------------------------------------------------------------------------------------
namespace {
// persistent state: everything except per-step temporaries (events, island arrays)
const int kSnapshotFields[] = {
    B2GPU_F_GRAVITY, B2GPU_F_WORLD_FLAGS, B2GPU_F_INV_DT0, B2GPU_F_BODY_COUNT, B2GPU_F_PROXY_COUNT, B2GPU_F_CONTACT_COUNT,
    B2GPU_F_STATUS, B2GPU_F_STEP_COUNT, B2GPU_F_BODY_XF, B2GPU_F_BODY_SWEEP, B2GPU_F_BODY_SWEEP0, B2GPU_F_BODY_VEL,
    B2GPU_F_BODY_FORCE, B2GPU_F_BODY_MASS, B2GPU_F_BODY_DAMP, B2GPU_F_BODY_FLAGS, B2GPU_F_FIXTURE_INFO, B2GPU_F_FIXTURE_SHAPE,
    B2GPU_F_FIXTURE_MAT, B2GPU_F_FIXTURE_FILTER, B2GPU_F_PROXY_FAT, B2GPU_F_PROXY_INFO, B2GPU_F_PROXY_MOVED, B2GPU_F_SHAPE_WORDS,
    B2GPU_F_CHAIN_VERTS, B2GPU_F_CONTACT_A, B2GPU_F_CONTACT_K, B2GPU_F_CONTACT_S, B2GPU_F_CONTACT_P0, B2GPU_F_CONTACT_P1,
    B2GPU_F_CONTACT_N, B2GPU_F_JOINT_COUNT, B2GPU_F_JOINT_I, B2GPU_F_JOINT_F0, B2GPU_F_JOINT_F1, B2GPU_F_JOINT_F2,
    B2GPU_F_JOINT_IMPULSE, B2GPU_F_FIXTURE_COUNT, B2GPU_F_SHAPE_WORD_COUNT, B2GPU_F_CHAIN_VERT_COUNT, B2GPU_F_PROXY_FREE_COUNT,
    B2GPU_F_PROXY_FREE, B2GPU_F_NEW_FIXTURE,
};
struct SnapshotHeader {
    uint32_t magic, version;
    int32_t numWorlds, NB, NF, NP, NC, NS, NV, NJ;
    int32_t needInitialPairs, anyContinuous;
};
constexpr uint32_t kSnapshotMagic = 0x62326773u;  // "b2gs"
size_t snapshotFieldBytes(const b2gpu_batch* b, int id) {
    return (size_t)b->desc.numWorlds * b->fields[id].capacity * b->fields[id].components * 4;
}
}  // namespace

int64_t b2gpu_batch_snapshot_size(b2gpu_batch* b) {
    if (!b) return 0;
    size_t n = sizeof(SnapshotHeader);
    for (int id : kSnapshotFields) n += snapshotFieldBytes(b, id);
    return (int64_t)n;
}

int b2gpu_batch_snapshot_save(b2gpu_batch* b, void* blob, int64_t size) {
    if (!b || !blob || size < b2gpu_batch_snapshot_size(b)) return fail(B2GPU_ERR_INVALID, "snapshot buffer too small");
    SnapshotHeader h = {kSnapshotMagic, 1, b->desc.numWorlds, b->A.NB, b->A.NF, b->A.NP, b->A.NC, b->A.NS, b->A.NV, b->A.NJ,
                        b->needInitialPairs ? 1 : 0, b->anyContinuous ? 1 : 0};
    unsigned char* p = static_cast<unsigned char*>(blob);
    memcpy(p, &h, sizeof(h));
    p += sizeof(h);
    for (int id : kSnapshotFields) {
        int rc = transferField(b, id, p, 0, b->desc.numWorlds, false);
        if (rc != B2GPU_OK) return rc;
        p += snapshotFieldBytes(b, id);
    }
    return B2GPU_OK;
}

int b2gpu_batch_snapshot_load(b2gpu_batch* b, const void* blob, int64_t size) {
    if (!b || !blob || size < b2gpu_batch_snapshot_size(b)) return fail(B2GPU_ERR_INVALID, "snapshot blob too small");
    SnapshotHeader h;
    const unsigned char* p = static_cast<const unsigned char*>(blob);
    memcpy(&h, p, sizeof(h));
    if (h.magic != kSnapshotMagic || h.version != 1) return fail(B2GPU_ERR_INVALID, "not a b2gpu snapshot");
    if (h.numWorlds != b->desc.numWorlds || h.NB != b->A.NB || h.NF != b->A.NF || h.NP != b->A.NP || h.NC != b->A.NC ||
        h.NS != b->A.NS || h.NV != b->A.NV || h.NJ != b->A.NJ)
        return fail(B2GPU_ERR_INVALID, "snapshot was taken from a batch with other capacities");
    p += sizeof(h);
    for (int id : kSnapshotFields) {
        int rc = transferField(b, id, const_cast<unsigned char*>(p), 0, b->desc.numWorlds, true);
        if (rc != B2GPU_OK) return rc;
        p += snapshotFieldBytes(b, id);
    }
    // the islands are rebuilt (they are a function of the restored state), buffered moves are looked at by the next
    // step's pair search, and e_newFixture / the continuous-physics switch come back as they were
    CUDA_TRY(cudaMemsetAsync(b->A.topoDirty, 1, (size_t)b->A.numTiles * kTile * sizeof(int), b->stream));
    CUDA_TRY(cudaMemsetAsync(b->A.pendingMoves, 1, (size_t)b->A.numTiles * kTile * sizeof(int), b->stream));
    b->needInitialPairs = h.needInitialPairs != 0;
    if ((h.anyContinuous != 0) != b->anyContinuous) {
        b->anyContinuous = h.anyContinuous != 0;
        invalidateGraphs(b);
    }
    b->needProxyIndex = true;
    if (!b->maySleep) {  // the restored state may hold sleepers: step_io goes back to forces-first
        b->maySleep = true;
        invalidateGraphs(b);
    }
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_spawn_records(b2gpu_batch* b, const uint32_t* words, const int64_t* offsets, int32_t count, int32_t* bodyIndices) {
    if (!b || count < 0 || (count > 0 && (!words || !offsets || !bodyIndices))) return fail(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    for (int32_t i = 0; i < count; ++i) {
        const int w = (int)words[offsets[i]];
        if (w < 0 || w >= b->desc.numWorlds) return fail(B2GPU_ERR_INVALID, "spawn %d names world %d outside the batch", i, w);
        if (i > 0 && w < (int)words[offsets[i - 1]]) return fail(B2GPU_ERR_INVALID, "spawn records must be grouped by ascending world");
    }
    CUDA_TRY(cudaSetDevice(b->desc.device));
    DevBuf dw, doff, dout;
    int rc;
    const size_t nwords = (size_t)offsets[count];
    if ((rc = dw.alloc(nwords * 4)) != B2GPU_OK) return rc;
    if ((rc = doff.alloc((size_t)(count + 1) * sizeof(int64_t))) != B2GPU_OK) return rc;
    if ((rc = dout.alloc((size_t)count * sizeof(int32_t))) != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(dw.p, words, nwords * 4, cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaMemcpyAsync(doff.p, offsets, (size_t)(count + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream));
    k_spawn<<<(unsigned)((count + 127) / 128), 128, 0, b->stream>>>(b->A, (const uint32_t*)dw.p, (const long long*)doff.p, count, (int*)dout.p);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(bodyIndices, dout.p, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->needInitialPairs = true;  // some world has e_newFixture set: the next Step starts with FindNewContacts
    b->needProxyIndex = true;
    for (int32_t i = 0; i < count && !b->maySleep; ++i) {  // a body created asleep (b2BodyDef::awake == false)
        const uint32_t* r = words + offsets[i];
        const int flags = (int)r[6 + 24];
        if ((int)r[5] < 0 && (flags & kBodyTypeMask) == B2G_DYNAMIC_BODY && (flags & kBodyAwake) == 0) {
            b->maySleep = true;
            invalidateGraphs(b);
        }
    }
    return B2GPU_OK;
}

int b2gpu_batch_joint_records(b2gpu_batch* b, const uint32_t* words, int32_t count, int32_t* jointIds) {
    if (!b || count < 0 || (count > 0 && (!words || !jointIds))) return fail(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    for (int32_t i = 0; i < count; ++i) {
        const int w = (int)words[(size_t)i * 21];
        if (w < 0 || w >= b->desc.numWorlds) return fail(B2GPU_ERR_INVALID, "joint %d names world %d outside the batch", i, w);
        if (i > 0 && w < (int)words[(size_t)(i - 1) * 21]) return fail(B2GPU_ERR_INVALID, "joint records must be grouped by ascending world");
    }
    CUDA_TRY(cudaSetDevice(b->desc.device));
    DevBuf dw, dout;
    int rc;
    if ((rc = dw.alloc((size_t)count * 21 * 4)) != B2GPU_OK) return rc;
    if ((rc = dout.alloc((size_t)count * sizeof(int32_t))) != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(dw.p, words, (size_t)count * 21 * 4, cudaMemcpyHostToDevice, b->stream));
    k_create_joints<<<(unsigned)((count + 127) / 128), 128, 0, b->stream>>>(b->A, (const uint32_t*)dw.p, count, (int*)dout.p);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(jointIds, dout.p, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_read_events(b2gpu_batch* b, int64_t* offsets, b2gpu_event* events, int64_t capacity) {
    if (!b || !offsets || capacity < 0) return fail(B2GPU_ERR_INVALID, "bad arguments");
    CUDA_TRY(cudaSetDevice(b->desc.device));
    const int W = b->desc.numWorlds;
    std::vector<int32_t> counts((size_t)W);
    int rc = transferField(b, B2GPU_F_EVENT_COUNT, counts.data(), 0, W, false);
    if (rc != B2GPU_OK) return rc;
    offsets[0] = 0;
    for (int w = 0; w < W; ++w) {
        const int n = counts[w] < b->A.NE ? counts[w] : b->A.NE;  // an overflowed arena holds its first NE events
        offsets[w + 1] = offsets[w] + (n > 0 ? n : 0);
    }
    const int64_t total = offsets[W] < capacity ? offsets[W] : capacity;
    if (!events || total <= 0) return B2GPU_OK;
    DevBuf dOff, dEv;
    if ((rc = dOff.alloc((size_t)(W + 1) * sizeof(int64_t))) != B2GPU_OK) return rc;
    if ((rc = dEv.alloc((size_t)total * sizeof(b2gpu_event))) != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(dOff.p, offsets, (size_t)(W + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, b->stream));
    k_gather_events<<<(unsigned)W, 128, 0, b->stream>>>(b->A, (const long long*)dOff.p, (b2gpu_event*)dEv.p, (long long)total);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(events, dEv.p, (size_t)total * sizeof(b2gpu_event), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_body_commands(b2gpu_batch* b, const b2gpu_body_command* commands, int64_t count) {
    if (!b || count < 0 || (count > 0 && !commands)) return fail(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    CUDA_TRY(cudaSetDevice(b->desc.device));
    // group by world, keeping the caller's order inside a world (already grouped input is passed through)
    bool sorted = true;
    for (int64_t i = 0; i < count; ++i) {
        if (commands[i].world < 0 || commands[i].world >= b->desc.numWorlds)
            return fail(B2GPU_ERR_INVALID, "command %lld names world %d outside the batch", (long long)i, commands[i].world);
        if (i > 0 && commands[i].world < commands[i - 1].world) sorted = false;
    }
    std::vector<b2gpu_body_command> grouped;
    if (!sorted) {
        grouped.assign(commands, commands + count);
        std::stable_sort(grouped.begin(), grouped.end(),
                         [](const b2gpu_body_command& x, const b2gpu_body_command& y) { return x.world < y.world; });
        commands = grouped.data();
    }
    const size_t bytes = (size_t)count * sizeof(b2gpu_body_command);
    if (b->cmdBytes < bytes) {
        if (b->cmdDev) cudaFree(b->cmdDev);
        b->cmdDev = nullptr;
        b->cmdBytes = 0;
        CUDA_TRY(cudaMalloc((void**)&b->cmdDev, bytes * 2));
        b->cmdBytes = bytes * 2;
    }
    CUDA_TRY(cudaMemcpyAsync(b->cmdDev, commands, bytes, cudaMemcpyHostToDevice, b->stream));
    k_body_commands<<<(unsigned)((count + 127) / 128), 128, 0, b->stream>>>(b->A, b->cmdDev, (long long)count);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    // the host array (or its grouped copy) may die when we return
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    for (int64_t i = 0; i < count && !b->maySleep; ++i)
        if (commands[i].op == B2GPU_CMD_SET_AWAKE && commands[i].wake == 0) {  // b2Body::SetAwake(false): a sleeper from now on
            b->maySleep = true;
            invalidateGraphs(b);
        }
    for (int64_t i = 0; i < count; ++i)  // proxies came or went: the body -> proxies index is rebuilt before the next step
        if (commands[i].op == B2GPU_CMD_DESTROY_BODY || commands[i].op == B2GPU_CMD_SET_ACTIVE || commands[i].op == B2GPU_CMD_DESTROY_FIXTURE) {
            b->needProxyIndex = true;
            break;
        }
    return B2GPU_OK;
}

// ---- stateless device entry points ---------------------------------------------------------------
namespace {
int requireDevice(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B2GPU_ERR_NO_DEVICE, "no CUDA device available (%s); box2dpp_b200 has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return fail(B2GPU_ERR_INVALID, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    return B2GPU_OK;
}
}  // namespace

int b2gpu_batch_query_aabb(b2gpu_batch* b, const b2gpu_aabb_query* queries, int64_t count, int32_t maxHits, int32_t* counts,
                           int32_t* hits) {
    if (!b || count < 0 || maxHits < 0 || (count > 0 && (!queries || !counts || (maxHits > 0 && !hits))))
        return fail(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    for (int64_t i = 0; i < count; ++i)
        if (queries[i].world < 0 || queries[i].world >= b->desc.numWorlds)
            return fail(B2GPU_ERR_INVALID, "query %lld names world %d outside the batch", (long long)i, queries[i].world);
    CUDA_TRY(cudaSetDevice(b->desc.device));
    DevBuf dq, dc, dh;
    int rc;
    if ((rc = dq.alloc((size_t)count * sizeof(b2gpu_aabb_query))) != B2GPU_OK) return rc;
    if ((rc = dc.alloc((size_t)count * sizeof(int32_t))) != B2GPU_OK) return rc;
    if ((rc = dh.alloc((size_t)count * maxHits * sizeof(int32_t))) != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(dq.p, queries, (size_t)count * sizeof(b2gpu_aabb_query), cudaMemcpyHostToDevice, b->stream));
    k_query_aabb<<<(unsigned)((count + 127) / 128), 128, 0, b->stream>>>(b->A, (const b2gpu_aabb_query*)dq.p, (long long)count, maxHits,
                                                                         (int*)dc.p, (int*)dh.p);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(counts, dc.p, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    if (maxHits > 0)
        CUDA_TRY(cudaMemcpyAsync(hits, dh.p, (size_t)count * maxHits * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_batch_raycast_closest(b2gpu_batch* b, const b2gpu_ray* rays, int64_t count, b2gpu_ray_hit* hits) {
    if (!b || count < 0 || (count > 0 && (!rays || !hits))) return fail(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    for (int64_t i = 0; i < count; ++i)
        if (rays[i].world < 0 || rays[i].world >= b->desc.numWorlds)
            return fail(B2GPU_ERR_INVALID, "ray %lld names world %d outside the batch", (long long)i, rays[i].world);
    CUDA_TRY(cudaSetDevice(b->desc.device));
    DevBuf dr, dh;
    int rc;
    if ((rc = dr.alloc((size_t)count * sizeof(b2gpu_ray))) != B2GPU_OK) return rc;
    if ((rc = dh.alloc((size_t)count * sizeof(b2gpu_ray_hit))) != B2GPU_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(dr.p, rays, (size_t)count * sizeof(b2gpu_ray), cudaMemcpyHostToDevice, b->stream));
    k_raycast_closest<<<(unsigned)((count + 127) / 128), 128, 0, b->stream>>>(b->A, (const b2gpu_ray*)dr.p, (long long)count,
                                                                              (b2gpu_ray_hit*)dh.p);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hits, dh.p, (size_t)count * sizeof(b2gpu_ray_hit), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return B2GPU_OK;
}

int b2gpu_sincosf(int32_t device, const float* x, float* s, float* c, int64_t n) {
    if (!x || !s || !c || n < 0) return fail(B2GPU_ERR_INVALID, "bad arguments");
    int rc = requireDevice(device);
    if (rc != B2GPU_OK || n == 0) return rc;
    DevBuf dx, ds, dc;
    if ((rc = dx.alloc(n * 4)) || (rc = ds.alloc(n * 4)) || (rc = dc.alloc(n * 4))) return rc;
    CUDA_TRY(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
    k_sincosf<<<(unsigned)((n + 255) / 256), 256>>>((const float*)dx.p, (float*)ds.p, (float*)dc.p, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(s, ds.p, n * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(c, dc.p, n * 4, cudaMemcpyDeviceToHost));
    return B2GPU_OK;
}

int b2gpu_collide_pairs(int32_t device, int32_t kind, int32_t n, const float* shapeA, const float* xfA, const float* shapeB,
                        const float* xfB, int32_t* ints, float* floats) {
    if (!shapeA || !xfA || !shapeB || !xfB || !ints || !floats || n < 0 || kind < 0 || kind > B2GPU_COLLIDE_EDGE_POLYGON)
        return fail(B2GPU_ERR_INVALID, "bad arguments");
    int rc = requireDevice(device);
    if (rc != B2GPU_OK || n == 0) return rc;
    const size_t wa = (kind == B2GPU_COLLIDE_CIRCLES) ? 3 : (kind >= B2GPU_COLLIDE_EDGE_CIRCLE ? 10 : 35);
    const size_t wb = (kind == B2GPU_COLLIDE_POLYGONS || kind == B2GPU_COLLIDE_EDGE_POLYGON) ? 35 : 3;
    DevBuf da, db, dxa, dxb, di, df;
    if ((rc = da.alloc(n * wa * 4)) || (rc = db.alloc(n * wb * 4)) || (rc = dxa.alloc(n * 16)) || (rc = dxb.alloc(n * 16)) ||
        (rc = di.alloc(n * 16)) || (rc = df.alloc(n * 32)))
        return rc;
    CUDA_TRY(cudaMemcpy(da.p, shapeA, n * wa * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(db.p, shapeB, n * wb * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dxa.p, xfA, (size_t)n * 16, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dxb.p, xfB, (size_t)n * 16, cudaMemcpyHostToDevice));
    k_collide_pairs<<<(n + 127) / 128, 128>>>(kind, n, (const float*)da.p, (const float*)dxa.p, (const float*)db.p,
                                              (const float*)dxb.p, (int*)di.p, (float*)df.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(ints, di.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(floats, df.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
    return B2GPU_OK;
}

}  // extern "C"
