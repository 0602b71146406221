/* TEST INFRASTRUCTURE — CPU oracle, shapes and narrow phase.  See b2o.h.
 * Paths cited are relative to /root/reference/Regression/box2d-ref/Box2DRef (vanilla), which is
 * bit-identical to the fork's Box2D/ on these functions except where SURVEY.md F5 says otherwise
 * (edge/chain-vs-polygon: vanilla is followed).
 */
#include <string.h>

#include "b2o.h"

/* ---- shapes ----------------------------------------------------------------------------- */

/* b2PolygonShape::SetAsBox(hx, hy) — Collision/Shapes/b2PolygonShape.cpp:32-44 */
void b2o_polygon_set_box(b2o_shape* s, float hx, float hy) {
    s->type = B2G_SHAPE_POLYGON;
    s->radius = B2O_POLYGON_RADIUS;
    s->count = 4;
    s->verts[0] = V2(-hx, -hy);
    s->verts[1] = V2(hx, -hy);
    s->verts[2] = V2(hx, hy);
    s->verts[3] = V2(-hx, hy);
    s->normals[0] = V2(0.0f, -1.0f);
    s->normals[1] = V2(1.0f, 0.0f);
    s->normals[2] = V2(0.0f, 1.0f);
    s->normals[3] = V2(-1.0f, 0.0f);
    s->centroid = V2(0.0f, 0.0f);
}

/* b2PolygonShape::SetAsBox(hx, hy, center, angle) — b2PolygonShape.cpp:46-71 */
void b2o_polygon_set_box_offset(b2o_shape* s, float hx, float hy, v2 center, float angle) {
    b2o_polygon_set_box(s, hx, hy);
    s->centroid = center;
    xform xf;
    xf.p = center;
    xf.q = rot_set(angle);
    for (int i = 0; i < s->count; ++i) {
        s->verts[i] = xf_mul(xf, s->verts[i]);
        s->normals[i] = rot_mul(xf.q, s->normals[i]);
    }
}

/* ComputeCentroid — b2PolygonShape.cpp:78-121 */
static v2 compute_centroid(const v2* vs, int count) {
    v2 c = V2(0.0f, 0.0f);
    float area = 0.0f;
    v2 pRef = V2(0.0f, 0.0f);
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
        v2 p1 = pRef;
        v2 p2 = vs[i];
        v2 p3 = i + 1 < count ? vs[i + 1] : vs[0];
        v2 e1 = v2_sub(p2, p1);
        v2 e2 = v2_sub(p3, p1);
        float D = v2_cross(e1, e2);
        float triangleArea = 0.5f * D;
        area += triangleArea;
        /* c += triangleArea * inv3 * (p1 + p2 + p3) */
        c = v2_add(c, v2_scale(triangleArea * inv3, v2_add(v2_add(p1, p2), p3)));
    }
    c = v2_scale(1.0f / area, c);
    return c;
}

/* b2PolygonShape::Set — b2PolygonShape.cpp:122-245 (welding, gift-wrapping hull, normals, centroid) */
void b2o_polygon_set(b2o_shape* s, const v2* vertices, int count) {
    s->type = B2G_SHAPE_POLYGON;
    s->radius = B2O_POLYGON_RADIUS;
    if (count < 3) { b2o_polygon_set_box(s, 1.0f, 1.0f); return; }
    int n = count < B2O_MAX_POLY ? count : B2O_MAX_POLY;
    v2 ps[B2O_MAX_POLY];
    int tempCount = 0;
    for (int i = 0; i < n; ++i) {
        v2 v = vertices[i];
        int unique = 1;
        for (int j = 0; j < tempCount; ++j) {
            if (v2_dist_sq(v, ps[j]) < 0.5f * B2O_LINEAR_SLOP) { unique = 0; break; }
        }
        if (unique) ps[tempCount++] = v;
    }
    n = tempCount;
    if (n < 3) { b2o_polygon_set_box(s, 1.0f, 1.0f); return; }
    int i0 = 0;
    float x0 = ps[0].x;
    for (int i = 1; i < n; ++i) {
        float x = ps[i].x;
        if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
    }
    int hull[B2O_MAX_POLY];
    int m = 0;
    int ih = i0;
    for (;;) {
        hull[m] = ih;
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            v2 r = v2_sub(ps[ie], ps[hull[m]]);
            v2 v = v2_sub(ps[j], ps[hull[m]]);
            float c = v2_cross(r, v);
            if (c < 0.0f) ie = j;
            if (c == 0.0f && v2_len_sq(v) > v2_len_sq(r)) ie = j;
        }
        ++m;
        ih = ie;
        if (ie == i0) break;
    }
    if (m < 3) { b2o_polygon_set_box(s, 1.0f, 1.0f); return; }
    s->count = m;
    for (int i = 0; i < m; ++i) s->verts[i] = ps[hull[i]];
    for (int i = 0; i < m; ++i) {
        int i1 = i;
        int i2 = i + 1 < m ? i + 1 : 0;
        v2 edge = v2_sub(s->verts[i2], s->verts[i1]);
        s->normals[i] = v2_cross_vs(edge, 1.0f);
        v2_normalize(&s->normals[i]);
    }
    s->centroid = compute_centroid(s->verts, m);
}

/* ComputeMass: b2PolygonShape.cpp:354-437, b2CircleShape.cpp:96-103, b2EdgeShape.cpp:133-140,
 * b2ChainShape.cpp:187-194 */
void b2o_shape_mass(const b2o_shape* s, float density, b2o_massdata* md) {
    if (s->type == B2G_SHAPE_POLYGON) {
        v2 center = V2(0.0f, 0.0f);
        float area = 0.0f;
        float I = 0.0f;
        v2 sref = V2(0.0f, 0.0f);
        for (int i = 0; i < s->count; ++i) sref = v2_add(sref, s->verts[i]);
        sref = v2_scale(1.0f / s->count, sref);
        const float k_inv3 = 1.0f / 3.0f;
        for (int i = 0; i < s->count; ++i) {
            v2 e1 = v2_sub(s->verts[i], sref);
            v2 e2 = i + 1 < s->count ? v2_sub(s->verts[i + 1], sref) : v2_sub(s->verts[0], sref);
            float D = v2_cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            center = v2_add(center, v2_scale(triangleArea * k_inv3, v2_add(e1, e2)));
            float ex1 = e1.x, ey1 = e1.y;
            float ex2 = e2.x, ey2 = e2.y;
            float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
            float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
            I += (0.25f * k_inv3 * D) * (intx2 + inty2);
        }
        md->mass = density * area;
        center = v2_scale(1.0f / area, center);
        md->center = v2_add(center, sref);
        md->I = density * I;
        md->I += md->mass * (v2_dot(md->center, md->center) - v2_dot(center, center));
    } else if (s->type == B2G_SHAPE_CIRCLE) {
        md->mass = density * B2O_PI * s->radius * s->radius;
        md->center = s->p;
        md->I = md->mass * (0.5f * s->radius * s->radius + v2_dot(s->p, s->p));
    } else if (s->type == B2G_SHAPE_EDGE) {
        md->mass = 0.0f;
        md->center = v2_scale(0.5f, v2_add(s->v1, s->v2_));
        md->I = 0.0f;
    } else {
        md->mass = 0.0f;
        md->center = V2(0.0f, 0.0f);
        md->I = 0.0f;
    }
}

/* ComputeAABB: b2PolygonShape.cpp:335-352, b2CircleShape.cpp:85-94, b2EdgeShape.cpp:118-131,
 * b2ChainShape.cpp:169-185 (chain children add no radius) */
void b2o_shape_aabb(const b2o_shape* s, xform xf, int child, aabb* out) {
    if (s->type == B2G_SHAPE_POLYGON) {
        v2 lower = xf_mul(xf, s->verts[0]);
        v2 upper = lower;
        for (int i = 1; i < s->count; ++i) {
            v2 v = xf_mul(xf, s->verts[i]);
            lower = v2_min(lower, v);
            upper = v2_max(upper, v);
        }
        v2 r = V2(s->radius, s->radius);
        out->lo = v2_sub(lower, r);
        out->hi = v2_add(upper, r);
    } else if (s->type == B2G_SHAPE_CIRCLE) {
        v2 p = v2_add(xf.p, rot_mul(xf.q, s->p));
        out->lo = V2(p.x - s->radius, p.y - s->radius);
        out->hi = V2(p.x + s->radius, p.y + s->radius);
    } else if (s->type == B2G_SHAPE_EDGE) {
        v2 a = xf_mul(xf, s->v1);
        v2 b = xf_mul(xf, s->v2_);
        v2 lower = v2_min(a, b);
        v2 upper = v2_max(a, b);
        v2 r = V2(s->radius, s->radius);
        out->lo = v2_sub(lower, r);
        out->hi = v2_add(upper, r);
    } else {
        int i1 = child;
        int i2 = child + 1;
        if (i2 == s->chainCount) i2 = 0;
        v2 a = xf_mul(xf, s->chain[i1]);
        v2 b = xf_mul(xf, s->chain[i2]);
        out->lo = v2_min(a, b);
        out->hi = v2_max(a, b);
    }
}

/* b2ChainShape::GetChildEdge — b2ChainShape.cpp:111-140 */
void b2o_chain_child_edge(const b2o_shape* c, int index, b2o_shape* e) {
    memset(e, 0, sizeof(*e));
    e->type = B2G_SHAPE_EDGE;
    e->radius = c->radius;
    e->v1 = c->chain[index + 0];
    e->v2_ = c->chain[index + 1];
    if (index > 0) { e->v0 = c->chain[index - 1]; e->has0 = 1; }
    else { e->v0 = c->prev; e->has0 = c->hasPrev; }
    if (index < c->chainCount - 2) { e->v3 = c->chain[index + 2]; e->has3 = 1; }
    else { e->v3 = c->next; e->has3 = c->hasNext; }
}

/* ---- polygon vs polygon ------------------------------------------------------------------ */

typedef struct { v2 v; uint32_t key; } clipv;
#define KEY_IA(k) ((int)((k) & 0xff))
#define KEY_IB(k) ((int)(((k) >> 8) & 0xff))
#define KEY_TA(k) ((int)(((k) >> 16) & 0xff))
#define KEY_TB(k) ((int)(((k) >> 24) & 0xff))

/* b2FindMaxSeparation — Collision/b2CollidePolygon.cpp:28-66 */
static float find_max_separation(int* edgeIndex, const b2o_shape* poly1, xform xf1, const b2o_shape* poly2,
                                 xform xf2) {
    int count1 = poly1->count, count2 = poly2->count;
    xform xf = xf_mulT_xx(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -B2O_MAXFLOAT;
    for (int i = 0; i < count1; ++i) {
        v2 n = rot_mul(xf.q, poly1->normals[i]);
        v2 v1 = xf_mul(xf, poly1->verts[i]);
        float si = B2O_MAXFLOAT;
        for (int j = 0; j < count2; ++j) {
            float sij = v2_dot(n, v2_sub(poly2->verts[j], v1));
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

/* b2FindIncidentEdge — b2CollidePolygon.cpp:68-110 */
static void find_incident_edge(clipv c[2], const b2o_shape* poly1, xform xf1, int edge1, const b2o_shape* poly2,
                               xform xf2) {
    v2 normal1 = rot_mulT(xf2.q, rot_mul(xf1.q, poly1->normals[edge1]));
    int index = 0;
    float minDot = B2O_MAXFLOAT;
    for (int i = 0; i < poly2->count; ++i) {
        float dot = v2_dot(normal1, poly2->normals[i]);
        if (dot < minDot) { minDot = dot; index = i; }
    }
    int i1 = index;
    int i2 = i1 + 1 < poly2->count ? i1 + 1 : 0;
    c[0].v = xf_mul(xf2, poly2->verts[i1]);
    c[0].key = b2o_make_key(edge1, i1, B2O_CF_FACE, B2O_CF_VERTEX);
    c[1].v = xf_mul(xf2, poly2->verts[i2]);
    c[1].key = b2o_make_key(edge1, i2, B2O_CF_FACE, B2O_CF_VERTEX);
}

/* b2ClipSegmentToLine — Collision/b2Collision.cpp:198-234 */
static int clip_segment_to_line(clipv vOut[2], const clipv vIn[2], v2 normal, float offset, int vertexIndexA) {
    int numOut = 0;
    float distance0 = v2_dot(normal, vIn[0].v) - offset;
    float distance1 = v2_dot(normal, vIn[1].v) - offset;
    if (distance0 <= 0.0f) vOut[numOut++] = vIn[0];
    if (distance1 <= 0.0f) vOut[numOut++] = vIn[1];
    if (distance0 * distance1 < 0.0f) {
        float interp = distance0 / (distance0 - distance1);
        vOut[numOut].v = v2_add(vIn[0].v, v2_scale(interp, v2_sub(vIn[1].v, vIn[0].v)));
        vOut[numOut].key = b2o_make_key(vertexIndexA, KEY_IB(vIn[0].key), B2O_CF_VERTEX, B2O_CF_FACE);
        ++numOut;
    }
    return numOut;
}

/* b2CollidePolygons — b2CollidePolygon.cpp:119-242 */
void b2o_collide_polygons_m(b2o_manifold* manifold, const b2o_shape* polyA, xform xfA, const b2o_shape* polyB,
                            xform xfB) {
    manifold->pointCount = 0;
    float totalRadius = polyA->radius + polyB->radius;
    int edgeA = 0;
    float separationA = find_max_separation(&edgeA, polyA, xfA, polyB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = find_max_separation(&edgeB, polyB, xfB, polyA, xfA);
    if (separationB > totalRadius) return;

    const b2o_shape *poly1, *poly2;
    xform xf1, xf2;
    int edge1;
    int flip;
    const float k_tol = 0.1f * B2O_LINEAR_SLOP;
    if (separationB > separationA + k_tol) {
        poly1 = polyB; poly2 = polyA; xf1 = xfB; xf2 = xfA; edge1 = edgeB;
        manifold->type = B2O_MANIFOLD_FACE_B;
        flip = 1;
    } else {
        poly1 = polyA; poly2 = polyB; xf1 = xfA; xf2 = xfB; edge1 = edgeA;
        manifold->type = B2O_MANIFOLD_FACE_A;
        flip = 0;
    }
    clipv incidentEdge[2];
    find_incident_edge(incidentEdge, poly1, xf1, edge1, poly2, xf2);
    int count1 = poly1->count;
    int iv1 = edge1;
    int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    v2 v11 = poly1->verts[iv1];
    v2 v12 = poly1->verts[iv2];
    v2 localTangent = v2_sub(v12, v11);
    v2_normalize(&localTangent);
    v2 localNormal = v2_cross_vs(localTangent, 1.0f);
    v2 planePoint = v2_scale(0.5f, v2_add(v11, v12));
    v2 tangent = rot_mul(xf1.q, localTangent);
    v2 normal = v2_cross_vs(tangent, 1.0f);
    v11 = xf_mul(xf1, v11);
    v12 = xf_mul(xf1, v12);
    float frontOffset = v2_dot(normal, v11);
    float sideOffset1 = -v2_dot(tangent, v11) + totalRadius;
    float sideOffset2 = v2_dot(tangent, v12) + totalRadius;
    clipv clipPoints1[2], clipPoints2[2];
    int np = clip_segment_to_line(clipPoints1, incidentEdge, v2_neg(tangent), sideOffset1, iv1);
    if (np < 2) return;
    np = clip_segment_to_line(clipPoints2, clipPoints1, tangent, sideOffset2, iv2);
    if (np < 2) return;
    manifold->localNormal = localNormal;
    manifold->localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = v2_dot(normal, clipPoints2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            b2o_mpoint* cp = manifold->points + pointCount;
            cp->localPoint = xf_mulT(xf2, clipPoints2[i].v);
            cp->key = clipPoints2[i].key;
            if (flip) {
                uint32_t k = cp->key;
                cp->key = b2o_make_key(KEY_IB(k), KEY_IA(k), KEY_TB(k), KEY_TA(k));
            }
            ++pointCount;
        }
    }
    manifold->pointCount = pointCount;
}

/* ---- circles ----------------------------------------------------------------------------- */

/* b2CollideCircles — Collision/b2CollideCircle.cpp:25-50 */
void b2o_collide_circles_m(b2o_manifold* manifold, const b2o_shape* circleA, xform xfA, const b2o_shape* circleB,
                           xform xfB) {
    manifold->pointCount = 0;
    v2 pA = xf_mul(xfA, circleA->p);
    v2 pB = xf_mul(xfB, circleB->p);
    v2 d = v2_sub(pB, pA);
    float distSqr = v2_dot(d, d);
    float rA = circleA->radius, rB = circleB->radius;
    float radius = rA + rB;
    if (distSqr > radius * radius) return;
    manifold->type = B2O_MANIFOLD_CIRCLES;
    manifold->localPoint = circleA->p;
    manifold->localNormal = V2(0.0f, 0.0f);
    manifold->pointCount = 1;
    manifold->points[0].localPoint = circleB->p;
    manifold->points[0].key = 0;
}

/* b2CollidePolygonAndCircle — b2CollideCircle.cpp:52-153 */
void b2o_collide_polygon_circle_m(b2o_manifold* manifold, const b2o_shape* polygonA, xform xfA,
                                  const b2o_shape* circleB, xform xfB) {
    manifold->pointCount = 0;
    v2 c = xf_mul(xfB, circleB->p);
    v2 cLocal = xf_mulT(xfA, c);
    int normalIndex = 0;
    float separation = -B2O_MAXFLOAT;
    float radius = polygonA->radius + circleB->radius;
    int vertexCount = polygonA->count;
    const v2* vertices = polygonA->verts;
    const v2* normals = polygonA->normals;
    for (int i = 0; i < vertexCount; ++i) {
        float s = v2_dot(normals[i], v2_sub(cLocal, vertices[i]));
        if (s > radius) return;
        if (s > separation) { separation = s; normalIndex = i; }
    }
    int vertIndex1 = normalIndex;
    int vertIndex2 = vertIndex1 + 1 < vertexCount ? vertIndex1 + 1 : 0;
    v2 v1 = vertices[vertIndex1];
    v2 v2v = vertices[vertIndex2];
    if (separation < B2O_EPSILON) {
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_FACE_A;
        manifold->localNormal = normals[normalIndex];
        manifold->localPoint = v2_scale(0.5f, v2_add(v1, v2v));
        manifold->points[0].localPoint = circleB->p;
        manifold->points[0].key = 0;
        return;
    }
    float u1 = v2_dot(v2_sub(cLocal, v1), v2_sub(v2v, v1));
    float u2 = v2_dot(v2_sub(cLocal, v2v), v2_sub(v1, v2v));
    if (u1 <= 0.0f) {
        if (v2_dist_sq(cLocal, v1) > radius * radius) return;
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_FACE_A;
        manifold->localNormal = v2_sub(cLocal, v1);
        v2_normalize(&manifold->localNormal);
        manifold->localPoint = v1;
        manifold->points[0].localPoint = circleB->p;
        manifold->points[0].key = 0;
    } else if (u2 <= 0.0f) {
        if (v2_dist_sq(cLocal, v2v) > radius * radius) return;
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_FACE_A;
        manifold->localNormal = v2_sub(cLocal, v2v);
        v2_normalize(&manifold->localNormal);
        manifold->localPoint = v2v;
        manifold->points[0].localPoint = circleB->p;
        manifold->points[0].key = 0;
    } else {
        v2 faceCenter = v2_scale(0.5f, v2_add(v1, v2v));
        float sep = v2_dot(v2_sub(cLocal, faceCenter), normals[vertIndex1]);
        if (sep > radius) return;
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_FACE_A;
        manifold->localNormal = normals[vertIndex1];
        manifold->localPoint = faceCenter;
        manifold->points[0].localPoint = circleB->p;
        manifold->points[0].key = 0;
    }
}

/* ---- edge vs circle ------------------------------------------------------------------------ */

/* b2CollideEdgeAndCircle — Collision/b2CollideEdge.cpp:29-155 */
void b2o_collide_edge_circle_m(b2o_manifold* manifold, const b2o_shape* edgeA, xform xfA, const b2o_shape* circleB,
                               xform xfB) {
    manifold->pointCount = 0;
    v2 Q = xf_mulT(xfA, xf_mul(xfB, circleB->p));
    v2 A = edgeA->v1, B = edgeA->v2_;
    v2 e = v2_sub(B, A);
    float u = v2_dot(e, v2_sub(B, Q));
    float v = v2_dot(e, v2_sub(Q, A));
    float radius = edgeA->radius + circleB->radius;
    if (v <= 0.0f) {
        v2 P = A;
        v2 d = v2_sub(Q, P);
        float dd = v2_dot(d, d);
        if (dd > radius * radius) return;
        if (edgeA->has0) {
            v2 A1 = edgeA->v0;
            v2 B1 = A;
            v2 e1 = v2_sub(B1, A1);
            float u1 = v2_dot(e1, v2_sub(B1, Q));
            if (u1 > 0.0f) return;
        }
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_CIRCLES;
        manifold->localNormal = V2(0.0f, 0.0f);
        manifold->localPoint = P;
        manifold->points[0].key = b2o_make_key(0, 0, B2O_CF_VERTEX, B2O_CF_VERTEX);
        manifold->points[0].localPoint = circleB->p;
        return;
    }
    if (u <= 0.0f) {
        v2 P = B;
        v2 d = v2_sub(Q, P);
        float dd = v2_dot(d, d);
        if (dd > radius * radius) return;
        if (edgeA->has3) {
            v2 B2 = edgeA->v3;
            v2 A2 = B;
            v2 e2 = v2_sub(B2, A2);
            float v2s = v2_dot(e2, v2_sub(Q, A2));
            if (v2s > 0.0f) return;
        }
        manifold->pointCount = 1;
        manifold->type = B2O_MANIFOLD_CIRCLES;
        manifold->localNormal = V2(0.0f, 0.0f);
        manifold->localPoint = P;
        manifold->points[0].key = b2o_make_key(1, 0, B2O_CF_VERTEX, B2O_CF_VERTEX);
        manifold->points[0].localPoint = circleB->p;
        return;
    }
    float den = v2_dot(e, e);
    v2 P = v2_scale(1.0f / den, v2_add(v2_scale(u, A), v2_scale(v, B)));
    v2 d = v2_sub(Q, P);
    float dd = v2_dot(d, d);
    if (dd > radius * radius) return;
    v2 n = V2(-e.y, e.x);
    if (v2_dot(n, v2_sub(Q, A)) < 0.0f) n = V2(-n.x, -n.y);
    v2_normalize(&n);
    manifold->pointCount = 1;
    manifold->type = B2O_MANIFOLD_FACE_A;
    manifold->localNormal = n;
    manifold->localPoint = A;
    manifold->points[0].key = b2o_make_key(0, 0, B2O_CF_FACE, B2O_CF_VERTEX);
    manifold->points[0].localPoint = circleB->p;
}

/* ---- edge vs polygon (b2EPCollider) — b2CollideEdge.cpp:157-700 ----------------------------- */

enum { AXIS_UNKNOWN = 0, AXIS_EDGE_A = 1, AXIS_EDGE_B = 2 };
typedef struct { int type; int index; float separation; } epaxis;

void b2o_collide_edge_polygon_m(b2o_manifold* manifold, const b2o_shape* edgeA, xform xfA,
                                const b2o_shape* polygonB, xform xfB) {
    xform m_xf = xf_mulT_xx(xfA, xfB);
    v2 m_centroidB = xf_mul(m_xf, polygonB->centroid);
    v2 m_v0 = edgeA->v0, m_v1 = edgeA->v1, m_v2 = edgeA->v2_, m_v3 = edgeA->v3;
    int hasVertex0 = edgeA->has0, hasVertex3 = edgeA->has3;
    v2 m_normal0 = V2(0, 0), m_normal1, m_normal2 = V2(0, 0), m_normal, m_lowerLimit, m_upperLimit;
    int m_front;

    v2 edge1 = v2_sub(m_v2, m_v1);
    v2_normalize(&edge1);
    m_normal1 = V2(edge1.y, -edge1.x);
    float offset1 = v2_dot(m_normal1, v2_sub(m_centroidB, m_v1));
    float offset0 = 0.0f, offset2 = 0.0f;
    int convex1 = 0, convex2 = 0;
    if (hasVertex0) {
        v2 edge0 = v2_sub(m_v1, m_v0);
        v2_normalize(&edge0);
        m_normal0 = V2(edge0.y, -edge0.x);
        convex1 = v2_cross(edge0, edge1) >= 0.0f;
        offset0 = v2_dot(m_normal0, v2_sub(m_centroidB, m_v0));
    }
    if (hasVertex3) {
        v2 edge2 = v2_sub(m_v3, m_v2);
        v2_normalize(&edge2);
        m_normal2 = V2(edge2.y, -edge2.x);
        convex2 = v2_cross(edge1, edge2) > 0.0f;
        offset2 = v2_dot(m_normal2, v2_sub(m_centroidB, m_v2));
    }
    v2 n1 = m_normal1, nn1 = v2_neg(m_normal1), nn0 = v2_neg(m_normal0), nn2 = v2_neg(m_normal2);
    if (hasVertex0 && hasVertex3) {
        if (convex1 && convex2) {
            m_front = offset0 >= 0.0f || offset1 >= 0.0f || offset2 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = m_normal0; m_upperLimit = m_normal2; }
            else { m_normal = nn1; m_lowerLimit = nn1; m_upperLimit = nn1; }
        } else if (convex1) {
            m_front = offset0 >= 0.0f || (offset1 >= 0.0f && offset2 >= 0.0f);
            if (m_front) { m_normal = n1; m_lowerLimit = m_normal0; m_upperLimit = n1; }
            else { m_normal = nn1; m_lowerLimit = nn2; m_upperLimit = nn1; }
        } else if (convex2) {
            m_front = offset2 >= 0.0f || (offset0 >= 0.0f && offset1 >= 0.0f);
            if (m_front) { m_normal = n1; m_lowerLimit = n1; m_upperLimit = m_normal2; }
            else { m_normal = nn1; m_lowerLimit = nn1; m_upperLimit = nn0; }
        } else {
            m_front = offset0 >= 0.0f && offset1 >= 0.0f && offset2 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = n1; m_upperLimit = n1; }
            else { m_normal = nn1; m_lowerLimit = nn2; m_upperLimit = nn0; }
        }
    } else if (hasVertex0) {
        if (convex1) {
            m_front = offset0 >= 0.0f || offset1 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = m_normal0; m_upperLimit = nn1; }
            else { m_normal = nn1; m_lowerLimit = n1; m_upperLimit = nn1; }
        } else {
            m_front = offset0 >= 0.0f && offset1 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = n1; m_upperLimit = nn1; }
            else { m_normal = nn1; m_lowerLimit = n1; m_upperLimit = nn0; }
        }
    } else if (hasVertex3) {
        if (convex2) {
            m_front = offset1 >= 0.0f || offset2 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = nn1; m_upperLimit = m_normal2; }
            else { m_normal = nn1; m_lowerLimit = nn1; m_upperLimit = n1; }
        } else {
            m_front = offset1 >= 0.0f && offset2 >= 0.0f;
            if (m_front) { m_normal = n1; m_lowerLimit = nn1; m_upperLimit = n1; }
            else { m_normal = nn1; m_lowerLimit = nn2; m_upperLimit = n1; }
        }
    } else {
        m_front = offset1 >= 0.0f;
        if (m_front) { m_normal = n1; m_lowerLimit = nn1; m_upperLimit = nn1; }
        else { m_normal = nn1; m_lowerLimit = n1; m_upperLimit = n1; }
    }

    int countB = polygonB->count;
    v2 pbv[B2O_MAX_POLY], pbn[B2O_MAX_POLY];
    for (int i = 0; i < countB; ++i) {
        pbv[i] = xf_mul(m_xf, polygonB->verts[i]);
        pbn[i] = rot_mul(m_xf.q, polygonB->normals[i]);
    }
    float m_radius = 2.0f * B2O_POLYGON_RADIUS;
    manifold->pointCount = 0;

    /* ComputeEdgeSeparation — b2CollideEdge.cpp:632-649 */
    epaxis edgeAxis;
    edgeAxis.type = AXIS_EDGE_A;
    edgeAxis.index = m_front ? 0 : 1;
    edgeAxis.separation = FLT_MAX;
    for (int i = 0; i < countB; ++i) {
        float s = v2_dot(m_normal, v2_sub(pbv[i], m_v1));
        if (s < edgeAxis.separation) edgeAxis.separation = s;
    }
    if (edgeAxis.type == AXIS_UNKNOWN) return;
    if (edgeAxis.separation > m_radius) return;

    /* ComputePolygonSeparation — b2CollideEdge.cpp:651-702 */
    epaxis polygonAxis;
    polygonAxis.type = AXIS_UNKNOWN;
    polygonAxis.index = -1;
    polygonAxis.separation = -FLT_MAX;
    {
        v2 perp = V2(-m_normal.y, m_normal.x);
        for (int i = 0; i < countB; ++i) {
            v2 n = v2_neg(pbn[i]);
            float s1 = v2_dot(n, v2_sub(pbv[i], m_v1));
            float s2 = v2_dot(n, v2_sub(pbv[i], m_v2));
            float s = f_min(s1, s2);
            if (s > m_radius) {
                polygonAxis.type = AXIS_EDGE_B;
                polygonAxis.index = i;
                polygonAxis.separation = s;
                break;
            }
            if (v2_dot(n, perp) >= 0.0f) {
                if (v2_dot(v2_sub(n, m_upperLimit), m_normal) < -B2O_ANGULAR_SLOP) continue;
            } else {
                if (v2_dot(v2_sub(n, m_lowerLimit), m_normal) < -B2O_ANGULAR_SLOP) continue;
            }
            if (s > polygonAxis.separation) {
                polygonAxis.type = AXIS_EDGE_B;
                polygonAxis.index = i;
                polygonAxis.separation = s;
            }
        }
    }
    if (polygonAxis.type != AXIS_UNKNOWN && polygonAxis.separation > m_radius) return;

    const float k_relativeTol = 0.98f;
    const float k_absoluteTol = 0.001f;
    epaxis primaryAxis;
    if (polygonAxis.type == AXIS_UNKNOWN) primaryAxis = edgeAxis;
    else if (polygonAxis.separation > k_relativeTol * edgeAxis.separation + k_absoluteTol) primaryAxis = polygonAxis;
    else primaryAxis = edgeAxis;

    clipv ie[2];
    int rf_i1, rf_i2;
    v2 rf_v1, rf_v2, rf_normal, rf_sideNormal1, rf_sideNormal2;
    float rf_sideOffset1, rf_sideOffset2;
    if (primaryAxis.type == AXIS_EDGE_A) {
        manifold->type = B2O_MANIFOLD_FACE_A;
        int bestIndex = 0;
        float bestValue = v2_dot(m_normal, pbn[0]);
        for (int i = 1; i < countB; ++i) {
            float value = v2_dot(m_normal, pbn[i]);
            if (value < bestValue) { bestValue = value; bestIndex = i; }
        }
        int i1 = bestIndex;
        int i2 = i1 + 1 < countB ? i1 + 1 : 0;
        ie[0].v = pbv[i1];
        ie[0].key = b2o_make_key(0, i1, B2O_CF_FACE, B2O_CF_VERTEX);
        ie[1].v = pbv[i2];
        ie[1].key = b2o_make_key(0, i2, B2O_CF_FACE, B2O_CF_VERTEX);
        if (m_front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = m_v1; rf_v2 = m_v2; rf_normal = m_normal1; }
        else { rf_i1 = 1; rf_i2 = 0; rf_v1 = m_v2; rf_v2 = m_v1; rf_normal = v2_neg(m_normal1); }
    } else {
        manifold->type = B2O_MANIFOLD_FACE_B;
        ie[0].v = m_v1;
        ie[0].key = b2o_make_key(0, primaryAxis.index, B2O_CF_VERTEX, B2O_CF_FACE);
        ie[1].v = m_v2;
        ie[1].key = b2o_make_key(0, primaryAxis.index, B2O_CF_VERTEX, B2O_CF_FACE);
        rf_i1 = primaryAxis.index;
        rf_i2 = rf_i1 + 1 < countB ? rf_i1 + 1 : 0;
        rf_v1 = pbv[rf_i1];
        rf_v2 = pbv[rf_i2];
        rf_normal = pbn[rf_i1];
    }
    rf_sideNormal1 = V2(rf_normal.y, -rf_normal.x);
    rf_sideNormal2 = v2_neg(rf_sideNormal1);
    rf_sideOffset1 = v2_dot(rf_sideNormal1, rf_v1);
    rf_sideOffset2 = v2_dot(rf_sideNormal2, rf_v2);
    clipv clipPoints1[2], clipPoints2[2];
    int np = clip_segment_to_line(clipPoints1, ie, rf_sideNormal1, rf_sideOffset1, rf_i1);
    if (np < 2) return;
    np = clip_segment_to_line(clipPoints2, clipPoints1, rf_sideNormal2, rf_sideOffset2, rf_i2);
    if (np < 2) return;
    if (primaryAxis.type == AXIS_EDGE_A) {
        manifold->localNormal = rf_normal;
        manifold->localPoint = rf_v1;
    } else {
        manifold->localNormal = polygonB->normals[rf_i1];
        manifold->localPoint = polygonB->verts[rf_i1];
    }
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = v2_dot(rf_normal, v2_sub(clipPoints2[i].v, rf_v1));
        if (separation <= m_radius) {
            b2o_mpoint* cp = manifold->points + pointCount;
            if (primaryAxis.type == AXIS_EDGE_A) {
                cp->localPoint = xf_mulT(m_xf, clipPoints2[i].v);
                cp->key = clipPoints2[i].key;
            } else {
                uint32_t k = clipPoints2[i].key;
                cp->localPoint = clipPoints2[i].v;
                cp->key = b2o_make_key(KEY_IB(k), KEY_IA(k), KEY_TB(k), KEY_TA(k));
            }
            ++pointCount;
        }
    }
    manifold->pointCount = pointCount;
}
