// b2g_query.cuh — batched read-only queries over the arenas: b2World::QueryAABB and
// b2World::RayCast for many (world, query) pairs in one launch.
//
// Reference: b2World::QueryAABB / RayCast  Dynamics/b2World.cpp:965-1020 (wrappers :965-985, :987-1011),
//            b2DynamicTree::Query / RayCast  Collision/b2DynamicTree.h:218-303,
//            b2Fixture::RayCast Dynamics/b2Fixture.h:336-339 -> shape RayCast:
//            b2PolygonShape.cpp:265-333, b2CircleShape.cpp:54-83, b2EdgeShape.cpp:49-116, b2ChainShape.cpp:148-167.
//
// The tree only prunes: a leaf is visited iff its fat AABB passes the overlap (and, for rays, the separating
// axis) test, and the set of visited leaves does not depend on the tree's shape.  So one thread
// per query walks the world's proxies in proxy-id order and applies the leaf tests directly.  What differs from
// the reference is only the ORDER in which hits are reported (proxy id here, tree traversal there); the closest
// ray hit is order independent except for exact ties in the fraction.
//
// One thread per query; queries ordered world-fastest make neighbouring threads read neighbouring lanes of the
// interleaved arenas (coalesced).
#pragma once

#include "b2g_kernels.cuh"

namespace b2g {

// b2PolygonShape::RayCast — Collision/Shapes/b2PolygonShape.cpp:265-333
__device__ __forceinline__ bool RayCastPolygon(const DevPoly& poly, Xf xf, Vec2 ip1, Vec2 ip2, float maxFraction, float* fraction,
                                               Vec2* normal) {
    const Vec2 p1 = MulT(xf.q, ip1 - xf.p);
    const Vec2 p2 = MulT(xf.q, ip2 - xf.p);
    const Vec2 d = p2 - p1;
    float lower = 0.0f, upper = maxFraction;
    int index = -1;
    for (int i = 0; i < poly.count(); ++i) {
        const Vec2 n = poly.normal(i);
        const float numerator = Dot(n, poly.vert(i) - p1);
        const float denominator = Dot(n, d);
        if (denominator == 0.0f) {
            if (numerator < 0.0f) return false;
        } else {
            if (denominator < 0.0f && numerator < lower * denominator) {
                lower = Div(numerator, denominator);
                index = i;
            } else if (denominator > 0.0f && numerator < upper * denominator) {
                upper = Div(numerator, denominator);
            }
        }
        if (upper < lower) return false;
    }
    if (index >= 0) {
        *fraction = lower;
        *normal = Mul(xf.q, poly.normal(index));
        return true;
    }
    return false;
}

// b2CircleShape::RayCast — Collision/Shapes/b2CircleShape.cpp:54-83
__device__ __forceinline__ bool RayCastCircle(const CircleData& c, Xf xf, Vec2 ip1, Vec2 ip2, float maxFraction, float* fraction,
                                              Vec2* normal) {
    const Vec2 position = xf.p + Mul(xf.q, c.p);
    const Vec2 s = ip1 - position;
    const float b = Dot(s, s) - c.radius * c.radius;
    const Vec2 r = ip2 - ip1;
    const float cc = Dot(s, r);
    const float rr = Dot(r, r);
    const float sigma = cc * cc - rr * b;
    if (sigma < 0.0f || rr < kEpsilon) return false;
    float a = -(cc + Sqrt(sigma));
    if (0.0f <= a && a <= maxFraction * rr) {
        a = Div(a, rr);
        *fraction = a;
        Vec2 n = s + a * r;
        Normalize(n);
        *normal = n;
        return true;
    }
    return false;
}

// b2EdgeShape::RayCast — Collision/Shapes/b2EdgeShape.cpp:49-116 (also a chain's child edge, b2ChainShape.cpp:148-167)
__device__ __forceinline__ bool RayCastEdge(Vec2 v1, Vec2 v2, Xf xf, Vec2 ip1, Vec2 ip2, float maxFraction, float* fraction,
                                            Vec2* normal) {
    const Vec2 p1 = MulT(xf.q, ip1 - xf.p);
    const Vec2 p2 = MulT(xf.q, ip2 - xf.p);
    const Vec2 d = p2 - p1;
    const Vec2 e = v2 - v1;
    Vec2 n = V2(e.y, -e.x);
    Normalize(n);
    const float numerator = Dot(n, v1 - p1);
    const float denominator = Dot(n, d);
    if (denominator == 0.0f) return false;
    const float t = Div(numerator, denominator);
    if (t < 0.0f || maxFraction < t) return false;
    const Vec2 q = p1 + t * d;
    const Vec2 r = v2 - v1;
    const float rr = Dot(r, r);
    if (rr == 0.0f) return false;
    const float s = Div(Dot(q - v1, r), rr);
    if (s < 0.0f || 1.0f < s) return false;
    *fraction = t;
    *normal = numerator > 0.0f ? -Mul(xf.q, n) : Mul(xf.q, n);
    return true;
}

// b2Fixture::RayCast for proxy (fixture, child) of this lane's world
__device__ __forceinline__ bool RayCastProxy(const Arena& A, int tile, int lane, int fixture, int child, int fi, Vec2 p1, Vec2 p2,
                                             float maxFraction, float* fraction, Vec2* normal) {
    const int shapeType = (fi >> 16) & 3, body = fi & 0xffff;
    const float* w = A.shapeWords + ((size_t)tile * A.NS + col(A.fixtureShape, tile, A.NF, lane)[fixture]) * kTile + lane;
    const Xf xf = LoadXf(col(A.bodyXf, tile, A.NB, lane), body);
    if (shapeType == B2G_SHAPE_POLYGON) return RayCastPolygon(DevPoly(w), xf, p1, p2, maxFraction, fraction, normal);
    if (shapeType == B2G_SHAPE_CIRCLE) return RayCastCircle(LoadCircle(w), xf, p1, p2, maxFraction, fraction, normal);
    if (shapeType == B2G_SHAPE_EDGE) {
        const EdgeData e = LoadEdge(w);
        return RayCastEdge(e.v1, e.v2, xf, p1, p2, maxFraction, fraction, normal);
    }
    Col<float2> verts = col(A.chainVerts, tile, A.NV, lane);
    const int first = __float_as_int(w[1 * kTile]), count = __float_as_int(w[2 * kTile]);
    int i2 = child + 1;
    if (i2 == count) i2 = 0;
    const float2 a = verts[first + child], b = verts[first + i2];
    return RayCastEdge(V2(a.x, a.y), V2(b.x, b.y), xf, p1, p2, maxFraction, fraction, normal);
}

// b2World::QueryAABB: every proxy whose fat AABB overlaps the query box, ascending proxy id.
// hits[q * maxHits + i] = fixture | child << 16; counts[q] = number of overlaps (may exceed maxHits: truncated).
__global__ void k_query_aabb(Arena A, const b2gpu_aabb_query* __restrict__ queries, long long n, int maxHits,
                             int* __restrict__ counts, int* __restrict__ hits) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const b2gpu_aabb_query Q = queries[q];
    const int tile = Q.world >> 5, lane = Q.world & 31;
    const size_t ws = (size_t)tile * kTile + lane;
    const int np = A.proxyCount[ws];
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    const float4 box = make_float4(Q.lower[0], Q.lower[1], Q.upper[0], Q.upper[1]);
    int m = 0;
    for (int p = 0; p < np; ++p) {
        if (!FatOverlap(proxyFat[p], box)) continue;
        if (m < maxHits) hits[q * maxHits + m] = proxyInfo[p];
        ++m;
    }
    counts[q] = m;
}

// b2World::RayCast with the usual closest-hit callback (ReportFixture returns `fraction`, so the ray is clipped to the
// nearest hit so far).  out.fixture = -1 when nothing is hit.
__global__ void k_raycast_closest(Arena A, const b2gpu_ray* __restrict__ rays, long long n, b2gpu_ray_hit* __restrict__ out) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const b2gpu_ray R = rays[q];
    const int tile = R.world >> 5, lane = R.world & 31;
    const size_t ws = (size_t)tile * kTile + lane;
    const int np = A.proxyCount[ws];
    Col<float4> proxyFat = col(A.proxyFat, tile, A.NP, lane);
    Col<int> proxyInfo = col(A.proxyInfo, tile, A.NP, lane);
    Col<int> fixtureInfo = col(A.fixtureInfo, tile, A.NF, lane);
    const Vec2 p1 = V2(R.p1[0], R.p1[1]), p2 = V2(R.p2[0], R.p2[1]);
    // b2DynamicTree::RayCast's leaf tests (b2DynamicTree.h:243-291)
    Vec2 r = p2 - p1;
    Normalize(r);
    const Vec2 v = Cross(1.0f, r);
    const Vec2 absV = V2(Abs(v.x), Abs(v.y));
    float maxFraction = R.maxFraction;
    b2gpu_ray_hit best;
    best.fixture = -1; best.child = 0; best.fraction = maxFraction;
    best.point[0] = best.point[1] = best.normal[0] = best.normal[1] = 0.0f;
    for (int p = 0; p < np; ++p) {
        const float4 fat = proxyFat[p];
        {
            const Vec2 t = p1 + maxFraction * (p2 - p1);
            const Vec2 lo = Min(p1, t), hi = Max(p1, t);
            if (!FatOverlap(fat, make_float4(lo.x, lo.y, hi.x, hi.y))) continue;
            const Vec2 c = 0.5f * (V2(fat.x, fat.y) + V2(fat.z, fat.w));
            const Vec2 h = 0.5f * (V2(fat.z, fat.w) - V2(fat.x, fat.y));
            const float separation = Abs(Dot(v, p1 - c)) - Dot(absV, h);
            if (separation > 0.0f) continue;
        }
        const int pi = proxyInfo[p];
        const int fixture = pi & 0xffff, child = (pi >> 16) & 0xffff;
        float fraction;
        Vec2 normal;
        if (!RayCastProxy(A, tile, lane, fixture, child, fixtureInfo[fixture], p1, p2, maxFraction, &fraction, &normal)) continue;
        // closest-hit callback: an equal fraction reported later would replace the hit in the reference too, but in its
        // tree order; here the lowest proxy id wins an exact tie
        if (best.fixture >= 0 && !(fraction < best.fraction)) continue;
        const Vec2 point = (1.0f - fraction) * p1 + fraction * p2;
        best.fixture = fixture; best.child = child; best.fraction = fraction;
        best.point[0] = point.x; best.point[1] = point.y;
        best.normal[0] = normal.x; best.normal[1] = normal.y;
        maxFraction = fraction;
    }
    out[q] = best;
}

}  // namespace b2g
