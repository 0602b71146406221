// b2g_distance.cuh — b2Distance (GJK) and b2TestOverlap(shape, shape) on the device.
//
// Reference: Collision/b2Distance.cpp:28-604, b2Distance.h:103-141; b2TestOverlap Collision/b2Collision.cpp:237-255.
// Used by the sensor branch of b2Contact::Update (k_collide) and by b2TimeOfImpact (b2g_toi.cuh).
// Included from b2g_kernels.cuh after the shape accessors; not a standalone header.
#pragma once
// (textually included inside namespace b2g)

// ---- b2DistanceProxy (b2Distance.cpp:33-88) ------------------------------------------------------
struct DistanceProxy {
    const float* w;  // polygons: address of vertex 0 in the shape table of this lane (vertex i at w[2i*32], w[(2i+1)*32])
    Vec2 v0, v1;     // circles, edges, chain children: the one or two vertices themselves
    int count;
    float radius;
    bool table;
    __device__ __forceinline__ Vec2 vertex(int i) const {
        if (table) return V2(w[(size_t)(2 * i) * kTile], w[(size_t)(2 * i + 1) * kTile]);
        return i == 0 ? v0 : v1;
    }
    // b2DistanceProxy::GetSupport (b2Distance.h:109-124)
    __device__ __forceinline__ int support(Vec2 d) const {
        int bestIndex = 0;
        float bestValue = Dot(vertex(0), d);
        for (int i = 1; i < count; ++i) {
            float value = Dot(vertex(i), d);
            if (value > bestValue) {
                bestIndex = i;
                bestValue = value;
            }
        }
        return bestIndex;
    }
};

__device__ __forceinline__ DistanceProxy MakeDistanceProxy(const Arena& A, int tile, int lane, int shapeOff, int shapeType, int child) {
    const float* w = A.shapeWords + ((size_t)tile * A.NS + shapeOff) * kTile + lane;
    DistanceProxy p;
    p.w = w;
    p.v0 = p.v1 = V2(0.0f, 0.0f);
    p.table = false;
    if (shapeType == B2G_SHAPE_POLYGON) {
        p.count = __float_as_int(w[0]);
        p.radius = w[1 * kTile];
        p.w = w + 4 * kTile;
        p.table = true;
    } else if (shapeType == B2G_SHAPE_CIRCLE) {
        p.count = 1;
        p.radius = w[0];
        p.v0 = V2(w[1 * kTile], w[2 * kTile]);
    } else if (shapeType == B2G_SHAPE_EDGE) {
        p.count = 2;
        p.radius = w[0];
        p.v0 = V2(w[3 * kTile], w[4 * kTile]);
        p.v1 = V2(w[5 * kTile], w[6 * kTile]);
    } else {  // chain child: vertices index and index + 1 (wrapping to 0)
        Col<float2> verts = col(A.chainVerts, tile, A.NV, lane);
        const int first = __float_as_int(w[1 * kTile]);
        const int n = __float_as_int(w[2 * kTile]);
        p.count = 2;
        p.radius = w[0];
        float2 a = verts[first + child];
        float2 b = child + 1 < n ? verts[first + child + 1] : verts[first];
        p.v0 = V2(a.x, a.y);
        p.v1 = V2(b.x, b.y);
    }
    return p;
}

// ---- b2Distance (GJK) ----------------------------------------------------------------------------
struct SimplexCache {
    float metric;
    int count;
    int indexA[3], indexB[3];
};
struct SimplexVertex {
    Vec2 wA, wB, w;
    float a;
    int indexA, indexB;
};
struct Simplex {
    SimplexVertex v[3];
    int count;

    __device__ float metric() const {
        if (count == 2) return Length(v[0].w - v[1].w);
        if (count == 3) return Cross(v[1].w - v[0].w, v[2].w - v[0].w);
        return 0.0f;
    }
    __device__ void readCache(const SimplexCache* cache, const DistanceProxy& proxyA, Xf xfA, const DistanceProxy& proxyB, Xf xfB) {
        count = cache->count;
        for (int i = 0; i < count; ++i) {
            SimplexVertex* sv = v + i;
            sv->indexA = cache->indexA[i];
            sv->indexB = cache->indexB[i];
            sv->wA = Mul(xfA, proxyA.vertex(sv->indexA));
            sv->wB = Mul(xfB, proxyB.vertex(sv->indexB));
            sv->w = sv->wB - sv->wA;
            sv->a = 0.0f;
        }
        if (count > 1) {
            float metric1 = cache->metric;
            float metric2 = metric();
            if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < kEpsilon) count = 0;
        }
        if (count == 0) {
            SimplexVertex* sv = v;
            sv->indexA = 0;
            sv->indexB = 0;
            sv->wA = Mul(xfA, proxyA.vertex(0));
            sv->wB = Mul(xfB, proxyB.vertex(0));
            sv->w = sv->wB - sv->wA;
            sv->a = 1.0f;
            count = 1;
        }
    }
    __device__ void writeCache(SimplexCache* cache) const {
        cache->metric = metric();
        cache->count = count;
        for (int i = 0; i < count; ++i) {
            cache->indexA[i] = v[i].indexA;
            cache->indexB[i] = v[i].indexB;
        }
    }
    __device__ Vec2 searchDirection() const {
        if (count == 1) return -v[0].w;
        Vec2 e12 = v[1].w - v[0].w;
        float sgn = Cross(e12, -v[0].w);
        if (sgn > 0.0f) return Cross(1.0f, e12);  // origin is left of e12
        return Cross(e12, 1.0f);
    }
    __device__ void witnessPoints(Vec2* pA, Vec2* pB) const {
        if (count == 1) {
            *pA = v[0].wA;
            *pB = v[0].wB;
        } else if (count == 2) {
            *pA = v[0].a * v[0].wA + v[1].a * v[1].wA;
            *pB = v[0].a * v[0].wB + v[1].a * v[1].wB;
        } else {
            *pA = v[0].a * v[0].wA + v[1].a * v[1].wA + v[2].a * v[2].wA;
            *pB = *pA;
        }
    }
    // b2Simplex::Solve2 (b2Distance.cpp:250-281)
    __device__ void solve2() {
        Vec2 w1 = v[0].w, w2 = v[1].w;
        Vec2 e12 = w2 - w1;
        float d12_2 = -Dot(w1, e12);
        if (d12_2 <= 0.0f) {
            v[0].a = 1.0f;
            count = 1;
            return;
        }
        float d12_1 = Dot(w2, e12);
        if (d12_1 <= 0.0f) {
            v[1].a = 1.0f;
            count = 1;
            v[0] = v[1];
            return;
        }
        float inv_d12 = Div(1.0f, d12_1 + d12_2);
        v[0].a = d12_1 * inv_d12;
        v[1].a = d12_2 * inv_d12;
        count = 2;
    }
    // b2Simplex::Solve3 (b2Distance.cpp:288-398)
    __device__ void solve3() {
        Vec2 w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        Vec2 e12 = w2 - w1;
        float w1e12 = Dot(w1, e12), w2e12 = Dot(w2, e12);
        float d12_1 = w2e12, d12_2 = -w1e12;
        Vec2 e13 = w3 - w1;
        float w1e13 = Dot(w1, e13), w3e13 = Dot(w3, e13);
        float d13_1 = w3e13, d13_2 = -w1e13;
        Vec2 e23 = w3 - w2;
        float w2e23 = Dot(w2, e23), w3e23 = Dot(w3, e23);
        float d23_1 = w3e23, d23_2 = -w2e23;
        float n123 = Cross(e12, e13);
        float d123_1 = n123 * Cross(w2, w3);
        float d123_2 = n123 * Cross(w3, w1);
        float d123_3 = n123 * Cross(w1, w2);
        if (d12_2 <= 0.0f && d13_2 <= 0.0f) {
            v[0].a = 1.0f;
            count = 1;
            return;
        }
        if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
            float inv_d12 = Div(1.0f, d12_1 + d12_2);
            v[0].a = d12_1 * inv_d12;
            v[1].a = d12_2 * inv_d12;
            count = 2;
            return;
        }
        if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
            float inv_d13 = Div(1.0f, d13_1 + d13_2);
            v[0].a = d13_1 * inv_d13;
            v[2].a = d13_2 * inv_d13;
            count = 2;
            v[1] = v[2];
            return;
        }
        if (d12_1 <= 0.0f && d23_2 <= 0.0f) {
            v[1].a = 1.0f;
            count = 1;
            v[0] = v[1];
            return;
        }
        if (d13_1 <= 0.0f && d23_1 <= 0.0f) {
            v[2].a = 1.0f;
            count = 1;
            v[0] = v[2];
            return;
        }
        if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
            float inv_d23 = Div(1.0f, d23_1 + d23_2);
            v[1].a = d23_1 * inv_d23;
            v[2].a = d23_2 * inv_d23;
            count = 2;
            v[0] = v[2];
            return;
        }
        float inv_d123 = Div(1.0f, d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv_d123;
        v[1].a = d123_2 * inv_d123;
        v[2].a = d123_3 * inv_d123;
        count = 3;
    }
};

// b2Distance with useRadii = false (b2Distance.cpp:446-572); returns the distance, updates the cache.
__device__ __noinline__ float GjkDistance(SimplexCache* cache, const DistanceProxy* pA, Xf xfA, const DistanceProxy* pB, Xf xfB) {
    const DistanceProxy& proxyA = *pA;
    const DistanceProxy& proxyB = *pB;
    Simplex simplex;
    simplex.readCache(cache, proxyA, xfA, proxyB, xfB);
    const int k_maxIters = 20;
    int saveA[3], saveB[3];
    int saveCount = 0;
    int iter = 0;
    while (iter < k_maxIters) {
        saveCount = simplex.count;
        for (int i = 0; i < saveCount; ++i) {
            saveA[i] = simplex.v[i].indexA;
            saveB[i] = simplex.v[i].indexB;
        }
        if (simplex.count == 2) simplex.solve2();
        else if (simplex.count == 3) simplex.solve3();
        if (simplex.count == 3) break;
        Vec2 d = simplex.searchDirection();
        if (LengthSquared(d) < kEpsilon * kEpsilon) break;
        SimplexVertex* vertex = simplex.v + simplex.count;
        vertex->indexA = proxyA.support(MulT(xfA.q, -d));
        vertex->wA = Mul(xfA, proxyA.vertex(vertex->indexA));
        vertex->indexB = proxyB.support(MulT(xfB.q, d));
        vertex->wB = Mul(xfB, proxyB.vertex(vertex->indexB));
        vertex->w = vertex->wB - vertex->wA;
        ++iter;
        bool duplicate = false;
        for (int i = 0; i < saveCount; ++i) {
            if (vertex->indexA == saveA[i] && vertex->indexB == saveB[i]) {
                duplicate = true;
                break;
            }
        }
        if (duplicate) break;
        ++simplex.count;
    }
    Vec2 pointA, pointB;
    simplex.witnessPoints(&pointA, &pointB);
    float distance = Length(pointA - pointB);
    simplex.writeCache(cache);
    return distance;
}

// b2Distance's useRadii epilogue reduced to the distance it outputs (b2Distance.cpp:574-603)
__device__ __forceinline__ float ApplyRadii(float distance, float rA, float rB) {
    if (distance > rA + rB && distance > kEpsilon) return distance - (rA + rB);
    return 0.0f;
}

// b2TestOverlap(shapeA, indexA, shapeB, indexB, xfA, xfB) — b2Collision.cpp:237-255
__device__ __noinline__ bool TestOverlapShapes(const Arena* Ap, int tile, int lane, int shapeOffA, int typeA, int childA, Xf xfA,
                                               int shapeOffB, int typeB, int childB, Xf xfB) {
    const DistanceProxy pA = MakeDistanceProxy(*Ap, tile, lane, shapeOffA, typeA, childA);
    const DistanceProxy pB = MakeDistanceProxy(*Ap, tile, lane, shapeOffB, typeB, childB);
    SimplexCache cache;
    cache.count = 0;
    const float d = GjkDistance(&cache, &pA, xfA, &pB, xfB);
    return ApplyRadii(d, pA.radius, pB.radius) < 10.0f * kEpsilon;
}

