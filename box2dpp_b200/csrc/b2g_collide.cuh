// b2g_collide.cuh — narrow phase of the Step path for sm_100a: the five manifold functions the
// reference's contact classes dispatch to (Box2D/Dynamics/Contacts/b2Contact.cpp:42-76).
//
// Each function evaluates ONE contact and is called by one lane of the collide kernel
// (b2g_step.cu); shapes are read through small accessor types so the same code reads the
// world-interleaved shape table on the device and plain structs in the host unit tests.
//
// Reference (vanilla paths, Regression/box2d-ref/Box2DRef/Collision/; the fork's Box2D/Collision/
// is identical except for the edge/chain-vs-polygon defect SURVEY.md F5, where vanilla is followed):
//   b2CollidePolygons / b2FindMaxSeparation / b2FindIncidentEdge   b2CollidePolygon.cpp:28-245
//   b2ClipSegmentToLine                                            b2Collision.cpp:198-234
//   b2CollideCircles / b2CollidePolygonAndCircle                   b2CollideCircle.cpp:25-153
//   b2CollideEdgeAndCircle / b2EPCollider::Collide                 b2CollideEdge.cpp:29-700
#pragma once

#include "b2g_math.cuh"

namespace b2g {

enum ManifoldType { kManifoldCircles = 0, kManifoldFaceA = 1, kManifoldFaceB = 2 };
enum FeatureType { kFeatureVertex = 0, kFeatureFace = 1 };

// b2ContactID: key = indexA | indexB << 8 | typeA << 16 | typeB << 24 (Box2D/Collision/b2Collision.h:43-62)
B2G_HD uint32_t MakeKey(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 0xff) | ((uint32_t)(indexB & 0xff) << 8) | ((uint32_t)(typeA & 0xff) << 16) |
           ((uint32_t)(typeB & 0xff) << 24);
}
B2G_HD uint32_t FlipKey(uint32_t k) {
    return ((k >> 8) & 0xffu) | ((k & 0xffu) << 8) | (((k >> 24) & 0xffu) << 16) | (((k >> 16) & 0xffu) << 24);
}

// b2Manifold without the impulses (the caller carries those; b2Collision.h:74-112)
struct Manifold {
    Vec2 lp0, lp1;  // points[i].localPoint
    uint32_t key0, key1;
    Vec2 localNormal, localPoint;
    int type;
    int pointCount;
};

struct CircleData {
    Vec2 p;
    float radius;
};

// b2EdgeShape fields, also what b2ChainShape::GetChildEdge produces (b2ChainShape.cpp:111-140)
struct EdgeData {
    Vec2 v0, v1, v2, v3;
    bool has0, has3;
    float radius;
};

struct ClipVertex {
    Vec2 v;
    uint32_t key;
};

// b2ClipSegmentToLine.  The reference appends to vOut[numOut++]; here the two output slots are selected with the same
// predicates (both inputs kept leaves no room for the crossing point: d0, d1 <= 0 makes d0 * d1 >= 0), so the arrays are
// indexed statically and stay in registers instead of local memory.
B2G_HD int ClipSegmentToLine(ClipVertex vOut[2], const ClipVertex vIn[2], Vec2 normal, float offset, int vertexIndexA) {
    const float distance0 = Dot(normal, vIn[0].v) - offset;
    const float distance1 = Dot(normal, vIn[1].v) - offset;
    const bool keep0 = distance0 <= 0.0f, keep1 = distance1 <= 0.0f;
    const bool cross = distance0 * distance1 < 0.0f;
    ClipVertex x = vIn[1];  // the crossing point, evaluated only when it exists (as the reference does)
    if (cross) {
        float interp = Div(distance0, distance0 - distance1);
        x.v = vIn[0].v + interp * (vIn[1].v - vIn[0].v);
        x.key = MakeKey(vertexIndexA, (vIn[0].key >> 8) & 0xff, kFeatureVertex, kFeatureFace);
    }
    // slot 0: the first of (in0 if kept, in1 if kept, crossing point); slot 1: the next one
    const ClipVertex in0 = vIn[0], in1 = vIn[1];
    vOut[0].v.x = keep0 ? in0.v.x : (keep1 ? in1.v.x : x.v.x);
    vOut[0].v.y = keep0 ? in0.v.y : (keep1 ? in1.v.y : x.v.y);
    vOut[0].key = keep0 ? in0.key : (keep1 ? in1.key : x.key);
    const bool both = keep0 && keep1;
    vOut[1].v.x = both ? in1.v.x : x.v.x;
    vOut[1].v.y = both ? in1.v.y : x.v.y;
    vOut[1].key = both ? in1.key : x.key;
    return (keep0 ? 1 : 0) + (keep1 ? 1 : 0) + (cross ? 1 : 0);
}

// b2FindMaxSeparation
template <class P1, class P2>
B2G_HD float FindMaxSeparation(int* edgeIndex, const P1& poly1, Xf xf1, const P2& poly2, Xf xf2) {
    int count1 = poly1.count(), count2 = poly2.count();
    Xf xf = MulT(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -kMaxFloat;
#pragma unroll
    for (int i = 0; i < count1; ++i) {
        Vec2 n = Mul(xf.q, poly1.normal(i));
        Vec2 v1 = Mul(xf, poly1.vert(i));
        float si = kMaxFloat;
#pragma unroll
        for (int j = 0; j < count2; ++j) {
            float sij = Dot(n, poly2.vert(j) - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) {
            maxSeparation = si;
            bestIndex = i;
        }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

// b2FindIncidentEdge
template <class P1, class P2>
B2G_HD void FindIncidentEdge(ClipVertex c[2], const P1& poly1, Xf xf1, int edge1, const P2& poly2, Xf xf2) {
    Vec2 normal1 = MulT(xf2.q, Mul(xf1.q, poly1.normal(edge1)));
    int count2 = poly2.count();
    int index = 0;
    float minDot = kMaxFloat;
#pragma unroll
    for (int i = 0; i < count2; ++i) {
        float dot = Dot(normal1, poly2.normal(i));
        if (dot < minDot) {
            minDot = dot;
            index = i;
        }
    }
    int i1 = index;
    int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
    c[0].v = Mul(xf2, poly2.vert(i1));
    c[0].key = MakeKey(edge1, i1, kFeatureFace, kFeatureVertex);
    c[1].v = Mul(xf2, poly2.vert(i2));
    c[1].key = MakeKey(edge1, i2, kFeatureFace, kFeatureVertex);
}

// Second half of b2CollidePolygons once the reference polygon is chosen (b2CollidePolygon.cpp:162-242).
template <class P1, class P2>
B2G_HD void ClipPolygons(Manifold* m, const P1& poly1, Xf xf1, int edge1, const P2& poly2, Xf xf2, bool flip,
                         float totalRadius) {
    ClipVertex incidentEdge[2];
    FindIncidentEdge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
    int count1 = poly1.count();
    int iv1 = edge1;
    int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    Vec2 v11 = poly1.vert(iv1);
    Vec2 v12 = poly1.vert(iv2);
    Vec2 localTangent = v12 - v11;
    Normalize(localTangent);
    Vec2 localNormal = Cross(localTangent, 1.0f);
    Vec2 planePoint = 0.5f * (v11 + v12);
    Vec2 tangent = Mul(xf1.q, localTangent);
    Vec2 normal = Cross(tangent, 1.0f);
    v11 = Mul(xf1, v11);
    v12 = Mul(xf1, v12);
    float frontOffset = Dot(normal, v11);
    float sideOffset1 = -Dot(tangent, v11) + totalRadius;
    float sideOffset2 = Dot(tangent, v12) + totalRadius;
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = ClipSegmentToLine(clipPoints1, incidentEdge, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = ClipSegmentToLine(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    m->localNormal = localNormal;
    m->localPoint = planePoint;
    int pointCount = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float separation = Dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            Vec2 lp = MulT(xf2, clipPoints2[i].v);
            uint32_t key = flip ? FlipKey(clipPoints2[i].key) : clipPoints2[i].key;
            if (pointCount == 0) { m->lp0 = lp; m->key0 = key; }
            else { m->lp1 = lp; m->key1 = key; }
            ++pointCount;
        }
    }
    m->pointCount = pointCount;
}

// b2CollidePolygons
template <class PA, class PB>
B2G_HD void CollidePolygons(Manifold* m, const PA& polyA, Xf xfA, const PB& polyB, Xf xfB) {
    m->pointCount = 0;
    float totalRadius = polyA.radius() + polyB.radius();
    int edgeA = 0;
    float separationA = FindMaxSeparation(&edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = FindMaxSeparation(&edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;
    const float k_tol = 0.1f * kLinearSlop;
    if (separationB > separationA + k_tol) {
        m->type = kManifoldFaceB;
        ClipPolygons(m, polyB, xfB, edgeB, polyA, xfA, true, totalRadius);
    } else {
        m->type = kManifoldFaceA;
        ClipPolygons(m, polyA, xfA, edgeA, polyB, xfB, false, totalRadius);
    }
}

// b2CollideCircles
B2G_HD void CollideCircles(Manifold* m, CircleData circleA, Xf xfA, CircleData circleB, Xf xfB) {
    m->pointCount = 0;
    Vec2 pA = Mul(xfA, circleA.p);
    Vec2 pB = Mul(xfB, circleB.p);
    Vec2 d = pB - pA;
    float distSqr = Dot(d, d);
    float radius = circleA.radius + circleB.radius;
    if (distSqr > radius * radius) return;
    m->type = kManifoldCircles;
    m->localPoint = circleA.p;
    m->localNormal = V2(0.0f, 0.0f);
    m->pointCount = 1;
    m->lp0 = circleB.p;
    m->key0 = 0;
}

// b2CollidePolygonAndCircle
template <class PA>
B2G_HD void CollidePolygonAndCircle(Manifold* m, const PA& polygonA, Xf xfA, CircleData circleB, Xf xfB) {
    m->pointCount = 0;
    Vec2 c = Mul(xfB, circleB.p);
    Vec2 cLocal = MulT(xfA, c);
    int normalIndex = 0;
    float separation = -kMaxFloat;
    float radius = polygonA.radius() + circleB.radius;
    int vertexCount = polygonA.count();
    for (int i = 0; i < vertexCount; ++i) {
        float s = Dot(polygonA.normal(i), cLocal - polygonA.vert(i));
        if (s > radius) return;
        if (s > separation) {
            separation = s;
            normalIndex = i;
        }
    }
    int vertIndex1 = normalIndex;
    int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
    Vec2 v1 = polygonA.vert(vertIndex1);
    Vec2 v2 = polygonA.vert(vertIndex2);
    m->lp0 = circleB.p;
    m->key0 = 0;
    if (separation < kEpsilon) {
        m->pointCount = 1;
        m->type = kManifoldFaceA;
        m->localNormal = polygonA.normal(normalIndex);
        m->localPoint = 0.5f * (v1 + v2);
        return;
    }
    float u1 = Dot(cLocal - v1, v2 - v1);
    float u2 = Dot(cLocal - v2, v1 - v2);
    if (u1 <= 0.0f) {
        if (DistanceSquared(cLocal, v1) > radius * radius) return;
        m->pointCount = 1;
        m->type = kManifoldFaceA;
        m->localNormal = cLocal - v1;
        Normalize(m->localNormal);
        m->localPoint = v1;
    } else if (u2 <= 0.0f) {
        if (DistanceSquared(cLocal, v2) > radius * radius) return;
        m->pointCount = 1;
        m->type = kManifoldFaceA;
        m->localNormal = cLocal - v2;
        Normalize(m->localNormal);
        m->localPoint = v2;
    } else {
        Vec2 faceCenter = 0.5f * (v1 + v2);
        float sep = Dot(cLocal - faceCenter, polygonA.normal(vertIndex1));
        if (sep > radius) return;
        m->pointCount = 1;
        m->type = kManifoldFaceA;
        m->localNormal = polygonA.normal(vertIndex1);
        m->localPoint = faceCenter;
    }
}

// b2CollideEdgeAndCircle
B2G_HD void CollideEdgeAndCircle(Manifold* m, const EdgeData& edgeA, Xf xfA, CircleData circleB, Xf xfB) {
    m->pointCount = 0;
    Vec2 Q = MulT(xfA, Mul(xfB, circleB.p));
    Vec2 A = edgeA.v1, B = edgeA.v2;
    Vec2 e = B - A;
    float u = Dot(e, B - Q);
    float v = Dot(e, Q - A);
    float radius = edgeA.radius + circleB.radius;
    if (v <= 0.0f) {
        Vec2 P = A;
        Vec2 d = Q - P;
        float dd = Dot(d, d);
        if (dd > radius * radius) return;
        if (edgeA.has0) {
            Vec2 A1 = edgeA.v0;
            Vec2 B1 = A;
            Vec2 e1 = B1 - A1;
            float u1 = Dot(e1, B1 - Q);
            if (u1 > 0.0f) return;
        }
        m->pointCount = 1;
        m->type = kManifoldCircles;
        m->localNormal = V2(0.0f, 0.0f);
        m->localPoint = P;
        m->key0 = MakeKey(0, 0, kFeatureVertex, kFeatureVertex);
        m->lp0 = circleB.p;
        return;
    }
    if (u <= 0.0f) {
        Vec2 P = B;
        Vec2 d = Q - P;
        float dd = Dot(d, d);
        if (dd > radius * radius) return;
        if (edgeA.has3) {
            Vec2 B2 = edgeA.v3;
            Vec2 A2 = B;
            Vec2 e2 = B2 - A2;
            float v2 = Dot(e2, Q - A2);
            if (v2 > 0.0f) return;
        }
        m->pointCount = 1;
        m->type = kManifoldCircles;
        m->localNormal = V2(0.0f, 0.0f);
        m->localPoint = P;
        m->key0 = MakeKey(1, 0, kFeatureVertex, kFeatureVertex);
        m->lp0 = circleB.p;
        return;
    }
    float den = Dot(e, e);
    Vec2 P = Div(1.0f, den) * (u * A + v * B);
    Vec2 d = Q - P;
    float dd = Dot(d, d);
    if (dd > radius * radius) return;
    Vec2 n = V2(-e.y, e.x);
    if (Dot(n, Q - A) < 0.0f) n = V2(-n.x, -n.y);
    Normalize(n);
    m->pointCount = 1;
    m->type = kManifoldFaceA;
    m->localNormal = n;
    m->localPoint = A;
    m->key0 = MakeKey(0, 0, kFeatureFace, kFeatureVertex);
    m->lp0 = circleB.p;
}

// b2EPCollider::Collide (with ComputeEdgeSeparation / ComputePolygonSeparation inlined)
template <class PB>
B2G_HD void CollideEdgeAndPolygon(Manifold* m, const EdgeData& edgeA, Xf xfA, const PB& polygonB, Xf xfB) {
    Xf xf = MulT(xfA, xfB);
    Vec2 centroidB = Mul(xf, polygonB.centroid());
    Vec2 v0 = edgeA.v0, v1 = edgeA.v1, v2 = edgeA.v2, v3 = edgeA.v3;
    bool hasVertex0 = edgeA.has0, hasVertex3 = edgeA.has3;
    Vec2 normal0 = V2(0.0f, 0.0f), normal1, normal2 = V2(0.0f, 0.0f), normal, lowerLimit, upperLimit;
    bool front;

    Vec2 edge1 = v2 - v1;
    Normalize(edge1);
    normal1 = V2(edge1.y, -edge1.x);
    float offset1 = Dot(normal1, centroidB - v1);
    float offset0 = 0.0f, offset2 = 0.0f;
    bool convex1 = false, convex2 = false;
    if (hasVertex0) {
        Vec2 edge0 = v1 - v0;
        Normalize(edge0);
        normal0 = V2(edge0.y, -edge0.x);
        convex1 = Cross(edge0, edge1) >= 0.0f;
        offset0 = Dot(normal0, centroidB - v0);
    }
    if (hasVertex3) {
        Vec2 edge2 = v3 - v2;
        Normalize(edge2);
        normal2 = V2(edge2.y, -edge2.x);
        convex2 = Cross(edge1, edge2) > 0.0f;
        offset2 = Dot(normal2, centroidB - v2);
    }
    const Vec2 n1 = normal1, nn1 = -normal1, nn0 = -normal0, nn2 = -normal2;
    if (hasVertex0 && hasVertex3) {
        if (convex1 && convex2) {
            front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
            if (front) { normal = n1; lowerLimit = normal0; upperLimit = normal2; }
            else { normal = nn1; lowerLimit = nn1; upperLimit = nn1; }
        } else if (convex1) {
            front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
            if (front) { normal = n1; lowerLimit = normal0; upperLimit = n1; }
            else { normal = nn1; lowerLimit = nn2; upperLimit = nn1; }
        } else if (convex2) {
            front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
            if (front) { normal = n1; lowerLimit = n1; upperLimit = normal2; }
            else { normal = nn1; lowerLimit = nn1; upperLimit = nn0; }
        } else {
            front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
            if (front) { normal = n1; lowerLimit = n1; upperLimit = n1; }
            else { normal = nn1; lowerLimit = nn2; upperLimit = nn0; }
        }
    } else if (hasVertex0) {
        if (convex1) {
            front = offset0 >= 0.0f || offset1 >= 0.0f;
            if (front) { normal = n1; lowerLimit = normal0; upperLimit = nn1; }
            else { normal = nn1; lowerLimit = n1; upperLimit = nn1; }
        } else {
            front = offset0 >= 0.0f && offset1 >= 0.0f;
            if (front) { normal = n1; lowerLimit = n1; upperLimit = nn1; }
            else { normal = nn1; lowerLimit = n1; upperLimit = nn0; }
        }
    } else if (hasVertex3) {
        if (convex2) {
            front = offset1 >= 0.0f || offset2 >= 0.0f;
            if (front) { normal = n1; lowerLimit = nn1; upperLimit = normal2; }
            else { normal = nn1; lowerLimit = nn1; upperLimit = n1; }
        } else {
            front = offset1 >= 0.0f && offset2 >= 0.0f;
            if (front) { normal = n1; lowerLimit = nn1; upperLimit = n1; }
            else { normal = nn1; lowerLimit = nn2; upperLimit = n1; }
        }
    } else {
        front = offset1 >= 0.0f;
        if (front) { normal = n1; lowerLimit = nn1; upperLimit = nn1; }
        else { normal = nn1; lowerLimit = n1; upperLimit = n1; }
    }

    const int countB = polygonB.count();
    const float radius = 2.0f * kPolygonRadius;
    m->pointCount = 0;

    // ComputeEdgeSeparation (polygon B is transformed on the fly: m_polygonB.vertices[i] = b2Mul(m_xf, v_i))
    float edgeSeparation = kMaxFloat;
    for (int i = 0; i < countB; ++i) {
        float s = Dot(normal, Mul(xf, polygonB.vert(i)) - v1);
        if (s < edgeSeparation) edgeSeparation = s;
    }
    if (edgeSeparation > radius) return;

    // ComputePolygonSeparation
    int polyType = 0;  // 0 = e_unknown, 2 = e_edgeB
    int polyIndex = -1;
    float polySeparation = -kMaxFloat;
    {
        Vec2 perp = V2(-normal.y, normal.x);
        for (int i = 0; i < countB; ++i) {
            Vec2 n = -Mul(xf.q, polygonB.normal(i));
            Vec2 vi = Mul(xf, polygonB.vert(i));
            float s1 = Dot(n, vi - v1);
            float s2 = Dot(n, vi - v2);
            float s = Min(s1, s2);
            if (s > radius) {
                polyType = 2;
                polyIndex = i;
                polySeparation = s;
                break;
            }
            if (Dot(n, perp) >= 0.0f) {
                if (Dot(n - upperLimit, normal) < -kAngularSlop) continue;
            } else {
                if (Dot(n - lowerLimit, normal) < -kAngularSlop) continue;
            }
            if (s > polySeparation) {
                polyType = 2;
                polyIndex = i;
                polySeparation = s;
            }
        }
    }
    if (polyType != 0 && polySeparation > radius) return;

    const float k_relativeTol = 0.98f;
    const float k_absoluteTol = 0.001f;
    bool primaryIsEdgeA;
    if (polyType == 0) primaryIsEdgeA = true;
    else if (polySeparation > k_relativeTol * edgeSeparation + k_absoluteTol) primaryIsEdgeA = false;
    else primaryIsEdgeA = true;

    ClipVertex ie[2];
    int rf_i1, rf_i2;
    Vec2 rf_v1, rf_v2, rf_normal;
    if (primaryIsEdgeA) {
        m->type = kManifoldFaceA;
        int bestIndex = 0;
        float bestValue = Dot(normal, Mul(xf.q, polygonB.normal(0)));
        for (int i = 1; i < countB; ++i) {
            float value = Dot(normal, Mul(xf.q, polygonB.normal(i)));
            if (value < bestValue) {
                bestValue = value;
                bestIndex = i;
            }
        }
        int i1 = bestIndex;
        int i2 = i1 + 1 < countB ? i1 + 1 : 0;
        ie[0].v = Mul(xf, polygonB.vert(i1));
        ie[0].key = MakeKey(0, i1, kFeatureFace, kFeatureVertex);
        ie[1].v = Mul(xf, polygonB.vert(i2));
        ie[1].key = MakeKey(0, i2, kFeatureFace, kFeatureVertex);
        if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = v1; rf_v2 = v2; rf_normal = normal1; }
        else { rf_i1 = 1; rf_i2 = 0; rf_v1 = v2; rf_v2 = v1; rf_normal = -normal1; }
    } else {
        m->type = kManifoldFaceB;
        ie[0].v = v1;
        ie[0].key = MakeKey(0, polyIndex, kFeatureVertex, kFeatureFace);
        ie[1].v = v2;
        ie[1].key = MakeKey(0, polyIndex, kFeatureVertex, kFeatureFace);
        rf_i1 = polyIndex;
        rf_i2 = rf_i1 + 1 < countB ? rf_i1 + 1 : 0;
        rf_v1 = Mul(xf, polygonB.vert(rf_i1));
        rf_v2 = Mul(xf, polygonB.vert(rf_i2));
        rf_normal = Mul(xf.q, polygonB.normal(rf_i1));
    }
    Vec2 sideNormal1 = V2(rf_normal.y, -rf_normal.x);
    Vec2 sideNormal2 = -sideNormal1;
    float sideOffset1 = Dot(sideNormal1, rf_v1);
    float sideOffset2 = Dot(sideNormal2, rf_v2);
    ClipVertex clipPoints1[2], clipPoints2[2];
    int np = ClipSegmentToLine(clipPoints1, ie, sideNormal1, sideOffset1, rf_i1);
    if (np < 2) return;
    np = ClipSegmentToLine(clipPoints2, clipPoints1, sideNormal2, sideOffset2, rf_i2);
    if (np < 2) return;
    if (primaryIsEdgeA) {
        m->localNormal = rf_normal;
        m->localPoint = rf_v1;
    } else {
        m->localNormal = polygonB.normal(rf_i1);
        m->localPoint = polygonB.vert(rf_i1);
    }
    int pointCount = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float separation = Dot(rf_normal, clipPoints2[i].v - rf_v1);
        if (separation <= radius) {
            Vec2 lp;
            uint32_t key;
            if (primaryIsEdgeA) {
                lp = MulT(xf, clipPoints2[i].v);
                key = clipPoints2[i].key;
            } else {
                lp = clipPoints2[i].v;
                key = FlipKey(clipPoints2[i].key);
            }
            if (pointCount == 0) { m->lp0 = lp; m->key0 = key; }
            else { m->lp1 = lp; m->key1 = key; }
            ++pointCount;
        }
    }
    m->pointCount = pointCount;
}

}  // namespace b2g
