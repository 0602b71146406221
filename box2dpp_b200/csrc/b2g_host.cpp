// b2g_host.cpp — host layer of libb2gpu.so: turns the user's CreateBody / CreateFixture calls
// (b2g_scene, include/b2g_scene.h) into world-major arena fields and uploads them through the
// device-layer C ABI (b2gpu_batch_upload_field).  Everything the reference computes at scene-construction
// time is computed here with the same float32 expression order:
//   b2Body::b2Body                 Box2D/Dynamics/b2Body.cpp:27-108
//   b2Body::CreateFixture          Box2D/Dynamics/b2Body.cpp:168-214
//   b2Body::ResetMassData          Box2D/Dynamics/b2Body.cpp:283-352
//   b2Fixture::Create/CreateProxies Box2D/Dynamics/b2Fixture.cpp:43-69, 124-139
//   b2DynamicTree::CreateProxy     Box2D/Collision/b2DynamicTree.cpp:112-126 (fat AABB, node ids)
//   b2PolygonShape::Set/SetAsBox/ComputeMass/ComputeAABB  Box2D/Collision/Shapes/b2PolygonShape.cpp
//   b2CircleShape / b2EdgeShape / b2ChainShape ComputeMass/ComputeAABB
// Compiled with g++ -O2 -ffp-contract=off (no FMA contraction), never -ffast-math.
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/b2gpu.h"
#include "b2g_layout.h"
#include "b2g_math.cuh"

using namespace b2g;

extern "C" int b2gpu_host_set_error(int code, const char* msg);  // defined below; shares b2gpu_last_error()

namespace {

struct HostShape {
    int type = 0;
    float radius = 0.0f;
    int count = 0;
    Vec2 verts[kMaxPolygonVertices], normals[kMaxPolygonVertices], centroid;
    Vec2 p;                    // circle
    Vec2 v0, v1, v2, v3;       // edge
    bool has0 = false, has3 = false;
    std::vector<Vec2> chain;   // chain
    Vec2 prev, next;
    bool hasPrev = false, hasNext = false;
};

// b2PolygonShape::SetAsBox(hx, hy)
void SetAsBox(HostShape& s, float hx, float hy) {
    s.type = B2G_SHAPE_POLYGON;
    s.radius = kPolygonRadius;
    s.count = 4;
    s.verts[0] = V2(-hx, -hy);
    s.verts[1] = V2(hx, -hy);
    s.verts[2] = V2(hx, hy);
    s.verts[3] = V2(-hx, hy);
    s.normals[0] = V2(0.0f, -1.0f);
    s.normals[1] = V2(1.0f, 0.0f);
    s.normals[2] = V2(0.0f, 1.0f);
    s.normals[3] = V2(-1.0f, 0.0f);
    s.centroid = V2(0.0f, 0.0f);
}

// b2PolygonShape::SetAsBox(hx, hy, center, angle)
void SetAsBox(HostShape& s, float hx, float hy, Vec2 center, float angle) {
    SetAsBox(s, hx, hy);
    s.centroid = center;
    Xf xf;
    xf.p = center;
    xf.q = MakeRot(angle);
    for (int i = 0; i < s.count; ++i) {
        s.verts[i] = Mul(xf, s.verts[i]);
        s.normals[i] = Mul(xf.q, s.normals[i]);
    }
}

// ComputeCentroid (b2PolygonShape.cpp:73-121)
Vec2 ComputeCentroid(const Vec2* vs, int count) {
    Vec2 c = V2(0.0f, 0.0f);
    float area = 0.0f;
    Vec2 pRef = V2(0.0f, 0.0f);
    const float inv3 = 1.0f / 3.0f;
    for (int i = 0; i < count; ++i) {
        Vec2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
        Vec2 e1 = p2 - p1, e2 = p3 - p1;
        float D = Cross(e1, e2);
        float triangleArea = 0.5f * D;
        area += triangleArea;
        c = c + (triangleArea * inv3) * (p1 + p2 + p3);
    }
    c = (1.0f / area) * c;
    return c;
}

// b2PolygonShape::Set: weld, gift-wrap hull, normals, centroid
void SetPolygon(HostShape& s, const Vec2* vertices, int count) {
    s.type = B2G_SHAPE_POLYGON;
    s.radius = kPolygonRadius;
    if (count < 3) { SetAsBox(s, 1.0f, 1.0f); return; }
    int n = count < kMaxPolygonVertices ? count : kMaxPolygonVertices;
    Vec2 ps[kMaxPolygonVertices];
    int tempCount = 0;
    for (int i = 0; i < n; ++i) {
        Vec2 v = vertices[i];
        bool unique = true;
        for (int j = 0; j < tempCount; ++j) {
            if (DistanceSquared(v, ps[j]) < 0.5f * kLinearSlop) { unique = false; break; }
        }
        if (unique) ps[tempCount++] = v;
    }
    n = tempCount;
    if (n < 3) { SetAsBox(s, 1.0f, 1.0f); return; }
    int i0 = 0;
    float x0 = ps[0].x;
    for (int i = 1; i < n; ++i) {
        float x = ps[i].x;
        if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
    }
    int hull[kMaxPolygonVertices];
    int m = 0, ih = i0;
    for (;;) {
        hull[m] = ih;
        int ie = 0;
        for (int j = 1; j < n; ++j) {
            if (ie == ih) { ie = j; continue; }
            Vec2 r = ps[ie] - ps[hull[m]];
            Vec2 v = ps[j] - ps[hull[m]];
            float c = Cross(r, v);
            if (c < 0.0f) ie = j;
            if (c == 0.0f && LengthSquared(v) > LengthSquared(r)) ie = j;
        }
        ++m;
        ih = ie;
        if (ie == i0) break;
    }
    if (m < 3) { SetAsBox(s, 1.0f, 1.0f); return; }
    s.count = m;
    for (int i = 0; i < m; ++i) s.verts[i] = ps[hull[i]];
    for (int i = 0; i < m; ++i) {
        int i2 = i + 1 < m ? i + 1 : 0;
        Vec2 edge = s.verts[i2] - s.verts[i];
        s.normals[i] = Cross(edge, 1.0f);
        Normalize(s.normals[i]);
    }
    s.centroid = ComputeCentroid(s.verts, m);
}

struct MassData {
    float mass;
    Vec2 center;
    float I;
};

// b2*Shape::ComputeMass
MassData ComputeMass(const HostShape& s, float density) {
    MassData md;
    if (s.type == B2G_SHAPE_POLYGON) {
        Vec2 center = V2(0.0f, 0.0f);
        float area = 0.0f, I = 0.0f;
        Vec2 sref = V2(0.0f, 0.0f);
        for (int i = 0; i < s.count; ++i) sref = sref + s.verts[i];
        sref = (1.0f / s.count) * sref;
        const float k_inv3 = 1.0f / 3.0f;
        for (int i = 0; i < s.count; ++i) {
            Vec2 e1 = s.verts[i] - sref;
            Vec2 e2 = i + 1 < s.count ? s.verts[i + 1] - sref : s.verts[0] - sref;
            float D = Cross(e1, e2);
            float triangleArea = 0.5f * D;
            area += triangleArea;
            center = center + (triangleArea * k_inv3) * (e1 + e2);
            float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
            float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2;
            float inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
            I += (0.25f * k_inv3 * D) * (intx2 + inty2);
        }
        md.mass = density * area;
        center = (1.0f / area) * center;
        md.center = center + sref;
        md.I = density * I;
        md.I += md.mass * (Dot(md.center, md.center) - Dot(center, center));
    } else if (s.type == B2G_SHAPE_CIRCLE) {
        md.mass = density * kPi * s.radius * s.radius;
        md.center = s.p;
        md.I = md.mass * (0.5f * s.radius * s.radius + Dot(s.p, s.p));
    } else if (s.type == B2G_SHAPE_EDGE) {
        md.mass = 0.0f;
        md.center = 0.5f * (s.v1 + s.v2);
        md.I = 0.0f;
    } else {
        md.mass = 0.0f;
        md.center = V2(0.0f, 0.0f);
        md.I = 0.0f;
    }
    return md;
}

// b2*Shape::ComputeAABB
void ComputeAABB(const HostShape& s, Xf xf, int child, Vec2* lo, Vec2* hi) {
    if (s.type == B2G_SHAPE_POLYGON) {
        Vec2 lower = Mul(xf, s.verts[0]);
        Vec2 upper = lower;
        for (int i = 1; i < s.count; ++i) {
            Vec2 v = Mul(xf, s.verts[i]);
            lower = Min(lower, v);
            upper = Max(upper, v);
        }
        Vec2 r = V2(s.radius, s.radius);
        *lo = lower - r;
        *hi = upper + r;
    } else if (s.type == B2G_SHAPE_CIRCLE) {
        Vec2 p = xf.p + Mul(xf.q, s.p);
        *lo = V2(p.x - s.radius, p.y - s.radius);
        *hi = V2(p.x + s.radius, p.y + s.radius);
    } else if (s.type == B2G_SHAPE_EDGE) {
        Vec2 a = Mul(xf, s.v1), b = Mul(xf, s.v2);
        Vec2 r = V2(s.radius, s.radius);
        *lo = Min(a, b) - r;
        *hi = Max(a, b) + r;
    } else {
        int i1 = child, i2 = child + 1;
        if (i2 == (int)s.chain.size()) i2 = 0;
        Vec2 a = Mul(xf, s.chain[i1]), b = Mul(xf, s.chain[i2]);
        *lo = Min(a, b);
        *hi = Max(a, b);
    }
}

int BuildShape(HostShape& sh, const b2g_fixture_def& d, const float* V) {
    switch (d.shapeType) {
        case B2G_SHAPE_CIRCLE:
            sh.type = B2G_SHAPE_CIRCLE;
            sh.radius = d.radius;
            sh.p = V2(d.p[0], d.p[1]);
            return 1;
        case B2G_SHAPE_EDGE:
            sh.type = B2G_SHAPE_EDGE;
            sh.radius = kPolygonRadius;
            sh.v1 = V2(V[2 * d.firstVertex], V[2 * d.firstVertex + 1]);
            sh.v2 = V2(V[2 * d.firstVertex + 2], V[2 * d.firstVertex + 3]);
            sh.has0 = d.hasVertex0 != 0;
            sh.has3 = d.hasVertex3 != 0;
            sh.v0 = V2(d.vertex0[0], d.vertex0[1]);
            sh.v3 = V2(d.vertex3[0], d.vertex3[1]);
            return 1;
        case B2G_SHAPE_POLYGON:
            if (d.shapeKind == B2G_POLY_BOX) SetAsBox(sh, d.p[0], d.p[1]);
            else if (d.shapeKind == B2G_POLY_BOX_OFFSET) SetAsBox(sh, d.p[0], d.p[1], V2(d.center[0], d.center[1]), d.angle);
            else {
                Vec2 pts[kMaxPolygonVertices];
                int n = d.vertexCount < kMaxPolygonVertices ? d.vertexCount : kMaxPolygonVertices;
                for (int k = 0; k < n; ++k) pts[k] = V2(V[2 * (d.firstVertex + k)], V[2 * (d.firstVertex + k) + 1]);
                SetPolygon(sh, pts, n);
            }
            return 1;
        case B2G_SHAPE_CHAIN: {
            sh.type = B2G_SHAPE_CHAIN;
            sh.radius = kPolygonRadius;
            for (int k = 0; k < d.vertexCount; ++k) sh.chain.push_back(V2(V[2 * (d.firstVertex + k)], V[2 * (d.firstVertex + k) + 1]));
            if (d.shapeKind == B2G_CHAIN_LOOP) {  // b2ChainShape::CreateLoop
                sh.chain.push_back(sh.chain[0]);
                sh.prev = sh.chain[sh.chain.size() - 2];
                sh.next = sh.chain[1];
                sh.hasPrev = sh.hasNext = true;
            } else {
                sh.hasPrev = d.hasVertex0 != 0;
                sh.hasNext = d.hasVertex3 != 0;
                sh.prev = V2(d.vertex0[0], d.vertex0[1]);
                sh.next = V2(d.vertex3[0], d.vertex3[1]);
            }
            return (int)sh.chain.size() - 1;
        }
    }
    return 0;
}

inline float IntBits(int v) { float f; memcpy(&f, &v, 4); return f; }

// One world flattened into world-major field arrays of capacity-sized rows.
// The constructor of every joint class that only copies its def (b2RevoluteJoint.cpp:47-64, b2PrismaticJoint.cpp:100-125,
// b2DistanceJoint.cpp:51-62, b2WheelJoint.cpp:44-72, b2WeldJoint.cpp:52-66, b2RopeJoint.cpp:39-52, b2FrictionJoint.cpp:46-61,
// b2MotorJoint.cpp:46-62, b2PulleyJoint.cpp:59-83) as the device's joint columns: I[4], F0[4], F1[4], F2[4] (bit patterns).
// Gear and mouse joints read body state when they are built and are laid out by their own code in Flatten.
static void PlainJointColumns(const b2g_joint_def& d, uint32_t cols[16]) {
    auto fb = [](float v) { uint32_t u; memcpy(&u, &v, 4); return u; };
    cols[0] = (uint32_t)(d.type | ((d.collideConnected ? 1 : 0) << 8) | ((d.enableLimit ? 1 : 0) << 9) | ((d.enableMotor ? 1 : 0) << 10));
    cols[1] = (uint32_t)(d.bodyA | (d.bodyB << 16));
    cols[2] = 0;  // m_limitState = e_inactiveLimit
    cols[3] = 0;
    cols[4] = fb(d.localAnchorA[0]); cols[5] = fb(d.localAnchorA[1]); cols[6] = fb(d.localAnchorB[0]); cols[7] = fb(d.localAnchorB[1]);
    float f1[4], f2[4];
    if (d.type == B2G_JOINT_DISTANCE) { f1[0] = d.length; f1[1] = d.frequencyHz; f1[2] = d.dampingRatio; f1[3] = 0.0f; }
    else if (d.type == B2G_JOINT_ROPE) { f1[0] = d.length; f1[1] = 0.0f; f1[2] = 0.0f; f1[3] = 0.0f; }
    else if (d.type == B2G_JOINT_FRICTION) { f1[0] = d.length; f1[1] = d.maxMotor; f1[2] = 0.0f; f1[3] = 0.0f; }
    else if (d.type == B2G_JOINT_PULLEY) { f1[0] = d.localAxisA[0]; f1[1] = d.localAxisA[1]; f1[2] = d.lower; f1[3] = d.upper; }
    else if (d.type == B2G_JOINT_MOTOR) { f1[0] = d.length; f1[1] = d.maxMotor; f1[2] = d.referenceAngle; f1[3] = d.frequencyHz; }
    else if (d.type == B2G_JOINT_WHEEL) { f1[0] = d.frequencyHz; f1[1] = d.dampingRatio; f1[2] = 0.0f; f1[3] = d.maxMotor; }
    else if (d.type == B2G_JOINT_WELD) { f1[0] = d.referenceAngle; f1[1] = d.frequencyHz; f1[2] = d.dampingRatio; f1[3] = 0.0f; }
    else { f1[0] = d.referenceAngle; f1[1] = d.lower; f1[2] = d.upper; f1[3] = d.maxMotor; }
    Vec2 axis = V2(d.localAxisA[0], d.localAxisA[1]);
    if (d.type == B2G_JOINT_PRISMATIC) Normalize(axis);  // m_localXAxisA.Normalize()
    if (d.type == B2G_JOINT_PULLEY) {  // ratio, m_constant = lengthA + ratio * lengthB (b2PulleyJoint.cpp:76)
        f2[0] = d.motorSpeed; f2[1] = d.length + d.motorSpeed * d.maxMotor; f2[2] = 0.0f; f2[3] = 0.0f;
    } else {
        f2[0] = d.motorSpeed; f2[1] = axis.x; f2[2] = axis.y; f2[3] = 0.0f;
    }
    for (int k = 0; k < 4; ++k) { cols[8 + k] = fb(f1[k]); cols[12 + k] = fb(f2[k]); }
}

struct FlatWorld {
    std::vector<std::vector<uint32_t>> f;  // per field id, capacity * components words
    int fixtures = 0, proxies = 0, shapeWords = 0, chainVerts = 0;  // slots used
};

struct Caps {
    int NB, NF, NP, NC, NS, NV, NE;
};

// Shape-table words of one shape (layout in b2g_layout.h).  Chains reference chainVerts at chainOffset.
std::vector<uint32_t> ShapeWords(const HostShape& sh, int chainOffset) {
    std::vector<uint32_t> words;
    auto pushf = [&](float v) { uint32_t u; memcpy(&u, &v, 4); words.push_back(u); };
    auto pushi = [&](int v) { words.push_back((uint32_t)v); };
    if (sh.type == B2G_SHAPE_POLYGON) {
        pushi(sh.count); pushf(sh.radius); pushf(sh.centroid.x); pushf(sh.centroid.y);
        for (int k = 0; k < sh.count; ++k) { pushf(sh.verts[k].x); pushf(sh.verts[k].y); }
        for (int k = 0; k < sh.count; ++k) { pushf(sh.normals[k].x); pushf(sh.normals[k].y); }
    } else if (sh.type == B2G_SHAPE_CIRCLE) {
        pushf(sh.radius); pushf(sh.p.x); pushf(sh.p.y);
    } else if (sh.type == B2G_SHAPE_EDGE) {
        pushf(sh.radius);
        pushf(sh.v0.x); pushf(sh.v0.y); pushf(sh.v1.x); pushf(sh.v1.y);
        pushf(sh.v2.x); pushf(sh.v2.y); pushf(sh.v3.x); pushf(sh.v3.y);
        pushi((sh.has0 ? 1 : 0) | (sh.has3 ? 2 : 0));
    } else {
        pushf(sh.radius); pushi(chainOffset); pushi((int)sh.chain.size());
        pushf(sh.prev.x); pushf(sh.prev.y); pushf(sh.next.x); pushf(sh.next.y);
        pushi((sh.hasPrev ? 1 : 0) | (sh.hasNext ? 2 : 0));
    }
    return words;
}

// Exact per-world sizes: bodies, fixtures, proxies, shape-table words after deduplication, chain vertices.
void CountScene(const b2g_scene& s, int* nb, int* nf, int* np, int* ns, int* nv) {
    *nb = s.bodyCount;
    *nf = s.fixtureCount;
    int proxies = 0, words = 0, verts = 0;
    std::map<std::vector<uint32_t>, int> dedup;
    for (int i = 0; i < s.fixtureCount; ++i) {
        HostShape sh;
        int children = BuildShape(sh, s.fixtures[i], s.vertices);
        bool active = s.fixtures[i].body >= 0 && s.fixtures[i].body < s.bodyCount &&
                      (s.bodies[s.fixtures[i].body].flags & B2G_BD_ACTIVE) != 0;
        if (active) proxies += children;
        std::vector<uint32_t> w = ShapeWords(sh, verts);
        if (dedup.find(w) == dedup.end()) {
            dedup[w] = words;
            words += (int)w.size();
        }
        verts += (int)sh.chain.size();
    }
    *np = proxies; *ns = words; *nv = verts;
}

template <class T>
void put(std::vector<uint32_t>& dst, size_t wordIndex, T v) {
    static_assert(sizeof(T) == 4, "32-bit fields only");
    memcpy(&dst[wordIndex], &v, 4);
}

int Flatten(const b2g_scene& s, const Caps& cap, const std::vector<std::pair<int, int>>& shapes /*cap, comps per field*/,
            FlatWorld& out, std::string& err) {
    int nb, nf, np, ns, nv;
    CountScene(s, &nb, &nf, &np, &ns, &nv);
    if (nb > cap.NB || nf > cap.NF || np > cap.NP || nv > cap.NV || ns > cap.NS) {
        err = "scene exceeds the batch capacities (bodies/fixtures/proxies/chain vertices)";
        return B2GPU_ERR_CAPACITY;
    }
    out.f.assign(B2GPU_F_COUNT, {});
    for (int id = 0; id < B2GPU_F_COUNT; ++id) out.f[id].assign((size_t)shapes[id].first * shapes[id].second, 0u);

    struct HBody {
        int type, flags;
        Xf xf;
        Vec2 localCenter, c0, c;
        float a0, a;
        Vec2 v;
        float w;
        float mass, invMass, I, invI;
    };
    std::vector<HBody> bodies(nb);
    for (int i = 0; i < nb; ++i) {
        const b2g_body_def& d = s.bodies[i];
        HBody& b = bodies[i];
        b.type = d.type;
        b.flags = d.type & kBodyTypeMask;
        if (d.flags & B2G_BD_BULLET) b.flags |= kBodyBullet;
        if (d.flags & B2G_BD_FIXED_ROTATION) b.flags |= kBodyFixedRotation;
        if (d.flags & B2G_BD_ALLOW_SLEEP) b.flags |= kBodyAutoSleep;
        if (d.flags & B2G_BD_AWAKE) b.flags |= kBodyAwake;
        if (d.flags & B2G_BD_ACTIVE) b.flags |= kBodyActive;
        b.xf.p = V2(d.position[0], d.position[1]);
        b.xf.q = MakeRot(d.angle);
        b.localCenter = V2(0.0f, 0.0f);
        b.c0 = b.c = b.xf.p;
        b.a0 = b.a = d.angle;
        b.v = V2(d.linearVelocity[0], d.linearVelocity[1]);
        b.w = d.angularVelocity;
        b.mass = b.invMass = d.type == B2G_DYNAMIC_BODY ? 1.0f : 0.0f;
        b.I = b.invI = 0.0f;
    }

    std::vector<HostShape> fshapes(nf);
    std::map<std::vector<uint32_t>, int> shapeDedup;
    int shapeWordsUsed = 0, chainUsed = 0, proxyCount = 0;
    auto& SW = out.f[B2GPU_F_SHAPE_WORDS];
    auto& CV = out.f[B2GPU_F_CHAIN_VERTS];
    for (int i = 0; i < nf; ++i) {
        const b2g_fixture_def& d = s.fixtures[i];
        if (d.body < 0 || d.body >= nb) { err = "fixture references a body that does not exist"; return B2GPU_ERR_INVALID; }
        HostShape& sh = fshapes[i];
        int childCount = BuildShape(sh, d, s.vertices);
        // shape table entry (deduplicated within the world)
        std::vector<uint32_t> words = ShapeWords(sh, chainUsed);
        for (const Vec2& v : sh.chain) {
            put(CV, (size_t)chainUsed * 2, v.x);
            put(CV, (size_t)chainUsed * 2 + 1, v.y);
            ++chainUsed;
        }
        int offset;
        auto it = shapeDedup.find(words);
        if (it != shapeDedup.end()) {
            offset = it->second;
        } else {
            offset = shapeWordsUsed;
            if (offset + (int)words.size() > cap.NS) { err = "shape table capacity exceeded"; return B2GPU_ERR_CAPACITY; }
            for (size_t k = 0; k < words.size(); ++k) SW[offset + k] = words[k];
            shapeWordsUsed += (int)words.size();
            shapeDedup[words] = offset;
        }
        put(out.f[B2GPU_F_FIXTURE_INFO], i, d.body | (sh.type << 16) | ((d.isSensor ? 1 : 0) << 18) | (d.density > 0.0f ? kFixtureHasDensity : 0));
        {   // b2Fixture::GetMassData, kept for b2Body::ResetMassData on the device (SetType, DestroyFixture, CreateFixture)
            MassData md = ComputeMass(sh, d.density);
            put(out.f[B2GPU_F_FIXTURE_MASS], (size_t)i * 4, md.mass); put(out.f[B2GPU_F_FIXTURE_MASS], (size_t)i * 4 + 1, md.center.x);
            put(out.f[B2GPU_F_FIXTURE_MASS], (size_t)i * 4 + 2, md.center.y); put(out.f[B2GPU_F_FIXTURE_MASS], (size_t)i * 4 + 3, md.I);
        }
        put(out.f[B2GPU_F_FIXTURE_SHAPE], i, offset);
        put(out.f[B2GPU_F_FIXTURE_MAT], (size_t)i * 2, d.friction);
        put(out.f[B2GPU_F_FIXTURE_MAT], (size_t)i * 2 + 1, d.restitution);
        put(out.f[B2GPU_F_FIXTURE_FILTER], (size_t)i * 2, (int)d.categoryBits | ((int)d.maskBits << 16));
        put(out.f[B2GPU_F_FIXTURE_FILTER], (size_t)i * 2 + 1, (int)d.groupIndex);

        HBody& b = bodies[d.body];
        if (b.flags & kBodyActive) {  // b2Fixture::CreateProxies
            for (int k = 0; k < childCount; ++k) {
                Vec2 lo, hi;
                ComputeAABB(sh, b.xf, k, &lo, &hi);
                Vec2 r = V2(kAabbExtension, kAabbExtension);
                lo = lo - r;
                hi = hi + r;
                size_t p = proxyCount++;
                put(out.f[B2GPU_F_PROXY_FAT], p * 4 + 0, lo.x);
                put(out.f[B2GPU_F_PROXY_FAT], p * 4 + 1, lo.y);
                put(out.f[B2GPU_F_PROXY_FAT], p * 4 + 2, hi.x);
                put(out.f[B2GPU_F_PROXY_FAT], p * 4 + 3, hi.y);
                put(out.f[B2GPU_F_PROXY_INFO], p, i | (k << 16));
                put(out.f[B2GPU_F_PROXY_MOVED], p, 1);  // b2BroadPhase::CreateProxy buffers a move
            }
        }
        if (d.density > 0.0f) {  // b2Body::ResetMassData
            b.mass = b.invMass = b.I = b.invI = 0.0f;
            b.localCenter = V2(0.0f, 0.0f);
            if (b.type == B2G_STATIC_BODY || b.type == B2G_KINEMATIC_BODY) {
                b.c0 = b.c = b.xf.p;
                b.a0 = b.a;
            } else {
                Vec2 localCenter = V2(0.0f, 0.0f);
                for (int f = i; f >= 0; --f) {  // fixture list is head-inserted: newest first
                    if (s.fixtures[f].body != d.body || s.fixtures[f].density == 0.0f) continue;
                    MassData md = ComputeMass(fshapes[f], s.fixtures[f].density);
                    b.mass += md.mass;
                    localCenter = localCenter + md.mass * md.center;
                    b.I += md.I;
                }
                if (b.mass > 0.0f) {
                    b.invMass = 1.0f / b.mass;
                    localCenter = b.invMass * localCenter;
                } else {
                    b.mass = 1.0f;
                    b.invMass = 1.0f;
                }
                if (b.I > 0.0f && (b.flags & kBodyFixedRotation) == 0) {
                    b.I -= b.mass * Dot(localCenter, localCenter);
                    b.invI = 1.0f / b.I;
                } else {
                    b.I = 0.0f;
                    b.invI = 0.0f;
                }
                Vec2 oldCenter = b.c;
                b.localCenter = localCenter;
                b.c0 = b.c = Mul(b.xf, b.localCenter);
                b.v = b.v + Cross(b.w, b.c - oldCenter);
            }
        }
    }

    for (int i = 0; i < nb; ++i) {
        const HBody& b = bodies[i];
        const b2g_body_def& d = s.bodies[i];
        size_t o = (size_t)i * 4;
        auto& XF = out.f[B2GPU_F_BODY_XF];
        put(XF, o, b.xf.p.x); put(XF, o + 1, b.xf.p.y); put(XF, o + 2, b.xf.q.s); put(XF, o + 3, b.xf.q.c);
        auto& S1 = out.f[B2GPU_F_BODY_SWEEP];
        put(S1, o, b.c.x); put(S1, o + 1, b.c.y); put(S1, o + 2, b.a); put(S1, o + 3, 0.0f);
        auto& S0 = out.f[B2GPU_F_BODY_SWEEP0];
        put(S0, o, b.c0.x); put(S0, o + 1, b.c0.y); put(S0, o + 2, b.a0); put(S0, o + 3, 0.0f);
        auto& VE = out.f[B2GPU_F_BODY_VEL];
        put(VE, o, b.v.x); put(VE, o + 1, b.v.y); put(VE, o + 2, b.w); put(VE, o + 3, 0.0f);
        auto& MS = out.f[B2GPU_F_BODY_MASS];
        put(MS, o, b.invMass); put(MS, o + 1, b.invI); put(MS, o + 2, b.localCenter.x); put(MS, o + 3, b.localCenter.y);
        auto& DP = out.f[B2GPU_F_BODY_DAMP];
        put(DP, o, d.linearDamping); put(DP, o + 1, d.angularDamping); put(DP, o + 2, d.gravityScale); put(DP, o + 3, b.mass);
        put(out.f[B2GPU_F_BODY_FLAGS], i, b.flags);
    }
    put(out.f[B2GPU_F_GRAVITY], 0, s.gravity[0]);
    put(out.f[B2GPU_F_GRAVITY], 1, s.gravity[1]);
    put(out.f[B2GPU_F_WORLD_FLAGS], 0, s.worldFlags);
    put(out.f[B2GPU_F_INV_DT0], 0, 0.0f);
    put(out.f[B2GPU_F_BODY_COUNT], 0, nb);
    put(out.f[B2GPU_F_PROXY_COUNT], 0, proxyCount);
    put(out.f[B2GPU_F_FIXTURE_COUNT], 0, nf);
    put(out.f[B2GPU_F_SHAPE_WORD_COUNT], 0, shapeWordsUsed);
    put(out.f[B2GPU_F_CHAIN_VERT_COUNT], 0, chainUsed);
    put(out.f[B2GPU_F_NEW_FIXTURE], 0, 1);  // CreateFixture raised e_newFixture
    out.fixtures = nf; out.proxies = proxyCount; out.shapeWords = shapeWordsUsed; out.chainVerts = chainUsed;

    // ---- b2World::CreateJoint calls (b2World.cpp:211-270); the joint constructors copy the def
    // (b2RevoluteJoint.cpp:47-64, b2PrismaticJoint.cpp:100-125, b2DistanceJoint.cpp:51-62) ----
    const int nj = s.jointCount > 0 ? s.jointCount : 0;
    if (nj > shapes[B2GPU_F_JOINT_I].first || (nj > 0 && !s.joints)) {
        err = "scene exceeds the batch's joint capacity";
        return B2GPU_ERR_CAPACITY;
    }
    for (int j = 0; j < nj; ++j) {
        const b2g_joint_def& d = s.joints[j];
        if (d.type != B2G_JOINT_REVOLUTE && d.type != B2G_JOINT_PRISMATIC && d.type != B2G_JOINT_DISTANCE &&
            d.type != B2G_JOINT_WHEEL && d.type != B2G_JOINT_WELD && d.type != B2G_JOINT_ROPE && d.type != B2G_JOINT_FRICTION &&
            d.type != B2G_JOINT_MOTOR && d.type != B2G_JOINT_PULLEY && d.type != B2G_JOINT_MOUSE && d.type != B2G_JOINT_GEAR) {
            err = "unknown joint type (b2JointType values 1-11 are built)";
            return B2GPU_ERR_UNSUPPORTED;
        }
        const size_t o = (size_t)j * 4;
        if (d.type == B2G_JOINT_GEAR) {
            // b2GearJoint::b2GearJoint (b2GearJoint.cpp:47-133): bodies, anchors, axes and reference angles are copied
            // from joint1 / joint2 (revolute or prismatic, created earlier), the constant from the bodies' poses now.
            const int j1 = d.enableLimit, j2 = d.enableMotor;
            if (j1 < 0 || j1 >= j || j2 < 0 || j2 >= j) { err = "gear joint: joint1 / joint2 must be joints created before it"; return B2GPU_ERR_INVALID; }
            int bodyOf[4];  // A, B, C, D
            float side[2][6] = {{0.0f}};
            int sideType[2];
            float coordinate[2];
            for (int k = 0; k < 2; ++k) {
                const b2g_joint_def& g = s.joints[k == 0 ? j1 : j2];
                if (g.type != B2G_JOINT_REVOLUTE && g.type != B2G_JOINT_PRISMATIC) { err = "gear joint: joint1 / joint2 must be revolute or prismatic"; return B2GPU_ERR_INVALID; }
                if (g.bodyA < 0 || g.bodyA >= nb || g.bodyB < 0 || g.bodyB >= nb) { err = "joint references a body that does not exist"; return B2GPU_ERR_INVALID; }
                sideType[k] = g.type;
                const HBody& ground = bodies[g.bodyA];  // m_bodyC / m_bodyD = joint->GetBodyA()
                const HBody& moving = bodies[g.bodyB];  // m_bodyA / m_bodyB = joint->GetBodyB()
                bodyOf[2 + k] = g.bodyA;
                bodyOf[k] = g.bodyB;
                if (g.type == B2G_JOINT_REVOLUTE) {
                    side[k][0] = g.referenceAngle;
                    coordinate[k] = moving.a - ground.a - g.referenceAngle;
                } else {
                    Vec2 axis = V2(g.localAxisA[0], g.localAxisA[1]);
                    Normalize(axis);  // the prismatic joint's m_localXAxisA
                    const Vec2 pG = V2(g.localAnchorA[0], g.localAnchorA[1]), lM = V2(g.localAnchorB[0], g.localAnchorB[1]);
                    const Vec2 pM = MulT(ground.xf.q, Mul(moving.xf.q, lM) + (moving.xf.p - ground.xf.p));
                    coordinate[k] = Dot(pM - pG, axis);
                    side[k][0] = pG.x; side[k][1] = pG.y; side[k][2] = lM.x; side[k][3] = lM.y; side[k][4] = axis.x; side[k][5] = axis.y;
                }
            }
            if (bodyOf[0] == bodyOf[1]) { err = "gear joint connects a body to itself"; return B2GPU_ERR_INVALID; }
            const float ratio = d.motorSpeed;
            const float constant = coordinate[0] + ratio * coordinate[1];
            put(out.f[B2GPU_F_JOINT_I], o, d.type | ((d.collideConnected ? 1 : 0) << 8) | (sideType[0] << 11) | (sideType[1] << 13) | (j << 16));
            put(out.f[B2GPU_F_JOINT_I], o + 1, bodyOf[0] | (bodyOf[1] << 16));
            put(out.f[B2GPU_F_JOINT_I], o + 2, 0);
            put(out.f[B2GPU_F_JOINT_I], o + 3, bodyOf[2] | (bodyOf[3] << 16));
            for (int k = 0; k < 4; ++k) put(out.f[B2GPU_F_JOINT_F0], o + k, side[0][k]);
            put(out.f[B2GPU_F_JOINT_F1], o, side[0][4]); put(out.f[B2GPU_F_JOINT_F1], o + 1, side[0][5]);
            put(out.f[B2GPU_F_JOINT_F1], o + 2, ratio); put(out.f[B2GPU_F_JOINT_F1], o + 3, constant);
            for (int k = 0; k < 4; ++k) put(out.f[B2GPU_F_JOINT_F2], o + k, side[1][k]);
            // the last two constants ride in the unused half of the impulse record (x = m_impulse)
            put(out.f[B2GPU_F_JOINT_IMPULSE], o, 0.0f); put(out.f[B2GPU_F_JOINT_IMPULSE], o + 1, 0.0f);
            put(out.f[B2GPU_F_JOINT_IMPULSE], o + 2, side[1][4]); put(out.f[B2GPU_F_JOINT_IMPULSE], o + 3, side[1][5]);
            continue;
        }
        if (d.bodyA < 0 || d.bodyA >= nb || d.bodyB < 0 || d.bodyB >= nb || d.bodyA == d.bodyB) {
            err = "joint references a body that does not exist (or connects a body to itself)";
            return B2GPU_ERR_INVALID;
        }
        if (d.type == B2G_JOINT_MOUSE) {
            // b2MouseJoint::b2MouseJoint (b2MouseJoint.cpp:33-52): m_localAnchorB = b2MulT(bodyB's transform, target);
            // b2Body::GetMass() of body B is read by InitVelocityConstraints (:113) — carried in the record
            const Vec2 target = V2(d.localAnchorA[0], d.localAnchorA[1]);
            const Vec2 lB = MulT(bodies[d.bodyB].xf, target);
            put(out.f[B2GPU_F_JOINT_I], o, d.type | ((d.collideConnected ? 1 : 0) << 8) | (j << 16));
            put(out.f[B2GPU_F_JOINT_I], o + 1, d.bodyA | (d.bodyB << 16));
            put(out.f[B2GPU_F_JOINT_I], o + 2, 0); put(out.f[B2GPU_F_JOINT_I], o + 3, 0);
            put(out.f[B2GPU_F_JOINT_F0], o, target.x); put(out.f[B2GPU_F_JOINT_F0], o + 1, target.y);
            put(out.f[B2GPU_F_JOINT_F0], o + 2, lB.x); put(out.f[B2GPU_F_JOINT_F0], o + 3, lB.y);
            put(out.f[B2GPU_F_JOINT_F1], o, d.length); put(out.f[B2GPU_F_JOINT_F1], o + 1, d.frequencyHz);
            put(out.f[B2GPU_F_JOINT_F1], o + 2, d.dampingRatio); put(out.f[B2GPU_F_JOINT_F1], o + 3, bodies[d.bodyB].mass);
            continue;
        }
        uint32_t cols[16];
        PlainJointColumns(d, cols);
        put(out.f[B2GPU_F_JOINT_I], o, (int)cols[0] | (j << 16));  // the joint's id = its creation number in the world
        for (int k = 1; k < 4; ++k) out.f[B2GPU_F_JOINT_I][o + k] = cols[k];
        for (int k = 0; k < 4; ++k) {
            out.f[B2GPU_F_JOINT_F0][o + k] = cols[4 + k];
            out.f[B2GPU_F_JOINT_F1][o + k] = cols[8 + k];
            out.f[B2GPU_F_JOINT_F2][o + k] = cols[12 + k];
        }
    }
    put(out.f[B2GPU_F_JOINT_SEQ], 0, nj);
    put(out.f[B2GPU_F_JOINT_COUNT], 0, nj);
    return B2GPU_OK;
}

// Fields written at scene-construction time, in upload order (the move buffer last: it arms FindNewContacts).
const int kBuildFields[] = {
    B2GPU_F_GRAVITY, B2GPU_F_WORLD_FLAGS, B2GPU_F_INV_DT0, B2GPU_F_BODY_COUNT, B2GPU_F_PROXY_COUNT,
    B2GPU_F_CONTACT_COUNT, B2GPU_F_STATUS, B2GPU_F_EVENT_COUNT, B2GPU_F_STEP_COUNT,
    B2GPU_F_BODY_XF, B2GPU_F_BODY_SWEEP, B2GPU_F_BODY_SWEEP0, B2GPU_F_BODY_VEL, B2GPU_F_BODY_FORCE, B2GPU_F_BODY_MASS,
    B2GPU_F_BODY_DAMP, B2GPU_F_BODY_FLAGS, B2GPU_F_FIXTURE_INFO, B2GPU_F_FIXTURE_SHAPE, B2GPU_F_FIXTURE_MAT,
    B2GPU_F_FIXTURE_FILTER, B2GPU_F_PROXY_FAT, B2GPU_F_PROXY_INFO, B2GPU_F_SHAPE_WORDS, B2GPU_F_CHAIN_VERTS,
    B2GPU_F_JOINT_COUNT, B2GPU_F_JOINT_I, B2GPU_F_JOINT_F0, B2GPU_F_JOINT_F1, B2GPU_F_JOINT_F2, B2GPU_F_JOINT_IMPULSE,
    B2GPU_F_FIXTURE_COUNT, B2GPU_F_SHAPE_WORD_COUNT, B2GPU_F_CHAIN_VERT_COUNT, B2GPU_F_PROXY_FREE_COUNT, B2GPU_F_NEW_FIXTURE,
    B2GPU_F_JOINT_SEQ, B2GPU_F_FIXTURE_MASS,
    B2GPU_F_PROXY_MOVED,
};

}  // namespace

// Checks a scene description at the C ABI boundary: every index the host layer is going to follow must be in range,
// every shape must have the vertices its class needs (b2EdgeShape 2, b2PolygonShape::Set 3..8, b2ChainShape::CreateChain
// >= 2, CreateLoop >= 3).  Returns B2GPU_OK or sets the error message.
static int ValidateFixtures(const b2g_fixture_def* fixtures, int fixtureCount, int bodyCount, const float* vertices, int vertexCount,
                            const char* what) {
    char msg[160];
    if (fixtureCount < 0 || bodyCount < 0 || vertexCount < 0 || (fixtureCount > 0 && !fixtures)) {
        snprintf(msg, sizeof(msg), "%s: negative count or null fixture array", what);
        return b2gpu_host_set_error(B2GPU_ERR_INVALID, msg);
    }
    for (int i = 0; i < fixtureCount; ++i) {
        const b2g_fixture_def& d = fixtures[i];
        int need = 0, least = 0;
        switch (d.shapeType) {
            case B2G_SHAPE_CIRCLE: break;
            case B2G_SHAPE_EDGE: need = 2; break;
            case B2G_SHAPE_POLYGON:
                if (d.shapeKind == B2G_POLY_SET) { need = d.vertexCount; least = 3; }
                break;
            case B2G_SHAPE_CHAIN: need = d.vertexCount; least = d.shapeKind == B2G_CHAIN_LOOP ? 3 : 2; break;
            default:
                snprintf(msg, sizeof(msg), "%s: fixture %d has unknown shape type %d", what, i, d.shapeType);
                return b2gpu_host_set_error(B2GPU_ERR_INVALID, msg);
        }
        if (d.body < 0 || d.body >= bodyCount) {
            snprintf(msg, sizeof(msg), "%s: fixture %d names body %d of %d", what, i, d.body, bodyCount);
            return b2gpu_host_set_error(B2GPU_ERR_INVALID, msg);
        }
        if (need < least || (d.shapeType == B2G_SHAPE_POLYGON && need > kMaxPolygonVertices)) {
            snprintf(msg, sizeof(msg), "%s: fixture %d has %d vertices (its shape class needs %d..)", what, i, need, least);
            return b2gpu_host_set_error(B2GPU_ERR_INVALID, msg);
        }
        if (need > 0 && (!vertices || d.firstVertex < 0 || (long long)d.firstVertex + need > (long long)vertexCount)) {
            snprintf(msg, sizeof(msg), "%s: fixture %d reads vertices [%d, %d) of %d", what, i, d.firstVertex, d.firstVertex + need,
                     vertexCount);
            return b2gpu_host_set_error(B2GPU_ERR_INVALID, msg);
        }
    }
    return B2GPU_OK;
}
static int ValidateScene(const b2g_scene& s, int index) {
    char what[48];
    snprintf(what, sizeof(what), "scene %d", index);
    if (s.bodyCount < 0 || (s.bodyCount > 0 && !s.bodies) || s.jointCount < 0 || (s.jointCount > 0 && !s.joints))
        return b2gpu_host_set_error(B2GPU_ERR_INVALID, "scene: negative count or null body / joint array");
    int rc = ValidateFixtures(s.fixtures, s.fixtureCount, s.bodyCount, s.vertices, s.vertexCount, what);
    if (rc != B2GPU_OK) return rc;
    for (int j = 0; j < s.jointCount; ++j)
        if (s.joints[j].bodyA < 0 || s.joints[j].bodyA >= s.bodyCount || s.joints[j].bodyB < 0 || s.joints[j].bodyB >= s.bodyCount)
            return b2gpu_host_set_error(B2GPU_ERR_INVALID, "scene: a joint names a body that does not exist");
    return B2GPU_OK;
}

extern "C" {

int b2gpu_scene_capacities(const b2g_scene* scenes, int32_t sceneCount, b2gpu_batch_desc* desc) {
    if (!scenes || !desc || sceneCount <= 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    for (int i = 0; i < sceneCount; ++i) {
        int rc = ValidateScene(scenes[i], i);
        if (rc != B2GPU_OK) return rc;
    }
    int NB = 1, NF = 1, NP = 1, NS = 4, NV = 1;
    for (int i = 0; i < sceneCount; ++i) {
        int nb, nf, np, ns, nv;
        CountScene(scenes[i], &nb, &nf, &np, &ns, &nv);
        NB = nb > NB ? nb : NB; NF = nf > NF ? nf : NF; NP = np > NP ? np : NP; NS = ns > NS ? ns : NS; NV = nv > NV ? nv : NV;
    }
    desc->maxBodies = NB;
    desc->maxFixtures = NF;
    desc->maxProxies = NP;
    desc->maxShapeWords = NS;
    desc->maxChainVerts = NV;
    int NJ = 0;
    for (int i = 0; i < sceneCount; ++i) NJ = scenes[i].jointCount > NJ ? scenes[i].jointCount : NJ;
    desc->maxJoints = NJ;
    // worlds that mix shape types (or have sensors) take the type-binned narrow phase of k_collide
    bool mixed = false;
    for (int i = 0; i < sceneCount && !mixed; ++i) {
        int types = 0;
        for (int f = 0; f < scenes[i].fixtureCount; ++f) {
            types |= 1 << (scenes[i].fixtures[f].shapeType & 3);
            if (scenes[i].fixtures[f].isSensor) mixed = true;
        }
        if (types & (types - 1)) mixed = true;
    }
    if (mixed) desc->reserved |= 4;
    if (desc->maxContacts <= 0) desc->maxContacts = 4 * NP + 32;   // pyramid-20 peaks at 604 for 211 proxies
    if (desc->maxEvents <= 0) desc->maxEvents = desc->eventMask ? 3 * desc->maxContacts : 1;
    return B2GPU_OK;
}

int b2gpu_batch_build_worlds(b2gpu_batch* batch, const b2g_scene* scenes, int32_t sceneCount, int32_t firstWorld,
                             int32_t replicate) {
    if (!batch || !scenes || sceneCount <= 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    for (int i = 0; i < sceneCount; ++i) {
        int rc = ValidateScene(scenes[i], i);
        if (rc != B2GPU_OK) return rc;
    }
    if (replicate < 1) replicate = 1;
    if (replicate > 1 && sceneCount != 1) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "replicate needs exactly one scene");
    std::vector<std::pair<int, int>> shapes(B2GPU_F_COUNT);
    for (int id = 0; id < B2GPU_F_COUNT; ++id) {
        int rc = b2gpu_batch_field_shape(batch, id, &shapes[id].first, &shapes[id].second);
        if (rc != B2GPU_OK) return rc;
    }
    Caps cap;
    cap.NB = shapes[B2GPU_F_BODY_XF].first;
    cap.NF = shapes[B2GPU_F_FIXTURE_INFO].first;
    cap.NP = shapes[B2GPU_F_PROXY_FAT].first;
    cap.NC = shapes[B2GPU_F_CONTACT_A].first;
    cap.NS = shapes[B2GPU_F_SHAPE_WORDS].first;
    cap.NV = shapes[B2GPU_F_CHAIN_VERTS].first;
    cap.NE = shapes[B2GPU_F_EVENT_A].first;

    const int total = sceneCount * replicate;
    const int chunk = 2048;  // worlds per upload
    std::string err;
    FlatWorld replicated;
    if (replicate > 1) {
        int rc = Flatten(scenes[0], cap, shapes, replicated, err);
        if (rc != B2GPU_OK) return b2gpu_host_set_error(rc, err.c_str());
    }
    std::vector<std::vector<uint32_t>> staging(B2GPU_F_COUNT);
    for (int w0 = 0; w0 < total; w0 += chunk) {
        int n = total - w0 < chunk ? total - w0 : chunk;
        for (int id : kBuildFields) staging[id].assign((size_t)n * shapes[id].first * shapes[id].second, 0u);
        for (int i = 0; i < n; ++i) {
            FlatWorld local;
            const FlatWorld* fw = &replicated;
            if (replicate == 1) {
                int rc = Flatten(scenes[w0 + i], cap, shapes, local, err);
                if (rc != B2GPU_OK) return b2gpu_host_set_error(rc, err.c_str());
                fw = &local;
            }
            for (int id : kBuildFields) {
                size_t words = (size_t)shapes[id].first * shapes[id].second;
                memcpy(&staging[id][(size_t)i * words], fw->f[id].data(), words * 4);
            }
        }
        for (int id : kBuildFields) {
            int rc = b2gpu_batch_upload_field(batch, id, staging[id].data(), firstWorld + w0, n);
            if (rc != B2GPU_OK) return rc;
        }
    }
    return B2GPU_OK;
}

// Spawn record (32-bit words), consumed by k_spawn (b2g_step.cu):
//   header   world, fixtures, proxies, shapeWords, chainVerts, target body (-1 = create a new body; else b2Body::CreateFixture
//            on that existing body: the body part is ignored and the proxies are made on the device under its transform)
//   body     xf[4] sweep[4] sweep0[4] vel[4] mass[4] damp[4] flags
//   fixture  info (body field = 0), shape offset (relative), friction, restitution, filter[2], mass data[4]   x fixtures
//   proxy    fat AABB[4], fixture (relative) | child << 16                                          x proxies
//   shape table words (chain entries hold a relative first vertex), chain vertices xy
static int SpawnOrAttach(b2gpu_batch* batch, const b2gpu_spawn* spawns, const int32_t* targets, int32_t count,
                         const b2g_fixture_def* fixtures, const float* vertices, int32_t* bodyIndices) {
    if (!batch || count < 0 || (count > 0 && (!spawns || !bodyIndices))) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    std::vector<std::pair<int, int>> shapes(B2GPU_F_COUNT);
    for (int id = 0; id < B2GPU_F_COUNT; ++id) {
        int rc = b2gpu_batch_field_shape(batch, id, &shapes[id].first, &shapes[id].second);
        if (rc != B2GPU_OK) return rc;
    }
    Caps cap;
    cap.NB = shapes[B2GPU_F_BODY_XF].first; cap.NF = shapes[B2GPU_F_FIXTURE_INFO].first; cap.NP = shapes[B2GPU_F_PROXY_FAT].first;
    cap.NC = shapes[B2GPU_F_CONTACT_A].first; cap.NS = shapes[B2GPU_F_SHAPE_WORDS].first; cap.NV = shapes[B2GPU_F_CHAIN_VERTS].first;
    cap.NE = shapes[B2GPU_F_EVENT_A].first;
    // spawns of one world keep their order: stable sort by world, results mapped back
    std::vector<int> order(count);
    for (int i = 0; i < count; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return spawns[a].world < spawns[b].world; });
    std::vector<uint32_t> words;
    std::vector<int64_t> offsets(count + 1);
    std::string err;
    for (int n = 0; n < count; ++n) {
        const b2gpu_spawn& sp = spawns[order[n]];
        if (sp.fixtureCount < 0 || (sp.fixtureCount > 0 && !fixtures)) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "spawn without fixtures array");
        std::vector<b2g_fixture_def> fds(fixtures + (sp.fixtureCount ? sp.firstFixture : 0), fixtures + (sp.fixtureCount ? sp.firstFixture + sp.fixtureCount : 0));
        for (b2g_fixture_def& fd : fds) fd.body = 0;
        b2g_scene mini;
        memset(&mini, 0, sizeof(mini));
        mini.worldFlags = B2G_WORLD_DEFAULT_FLAGS;
        mini.bodyCount = 1;
        mini.fixtureCount = (int)fds.size();
        mini.bodies = &sp.body;
        mini.fixtures = fds.data();
        mini.vertices = vertices;
        FlatWorld fw;
        int rc = Flatten(mini, cap, shapes, fw, err);
        if (rc != B2GPU_OK) return b2gpu_host_set_error(rc, err.c_str());
        offsets[n] = (int64_t)words.size();
        auto push = [&](uint32_t v) { words.push_back(v); };
        auto pushField = [&](int id, size_t first, size_t nwords) { for (size_t k = 0; k < nwords; ++k) push(fw.f[id][first + k]); };
        push((uint32_t)sp.world); push((uint32_t)fw.fixtures); push((uint32_t)fw.proxies); push((uint32_t)fw.shapeWords); push((uint32_t)fw.chainVerts);
        push((uint32_t)(targets ? targets[order[n]] : -1));
        pushField(B2GPU_F_BODY_XF, 0, 4); pushField(B2GPU_F_BODY_SWEEP, 0, 4); pushField(B2GPU_F_BODY_SWEEP0, 0, 4);
        pushField(B2GPU_F_BODY_VEL, 0, 4); pushField(B2GPU_F_BODY_MASS, 0, 4); pushField(B2GPU_F_BODY_DAMP, 0, 4);
        pushField(B2GPU_F_BODY_FLAGS, 0, 1);
        for (int f = 0; f < fw.fixtures; ++f) {
            pushField(B2GPU_F_FIXTURE_INFO, f, 1); pushField(B2GPU_F_FIXTURE_SHAPE, f, 1);
            pushField(B2GPU_F_FIXTURE_MAT, (size_t)f * 2, 2); pushField(B2GPU_F_FIXTURE_FILTER, (size_t)f * 2, 2);
            pushField(B2GPU_F_FIXTURE_MASS, (size_t)f * 4, 4);
        }
        for (int q = 0; q < fw.proxies; ++q) { pushField(B2GPU_F_PROXY_FAT, (size_t)q * 4, 4); pushField(B2GPU_F_PROXY_INFO, q, 1); }
        pushField(B2GPU_F_SHAPE_WORDS, 0, fw.shapeWords);
        pushField(B2GPU_F_CHAIN_VERTS, 0, (size_t)fw.chainVerts * 2);
    }
    offsets[count] = (int64_t)words.size();
    std::vector<int32_t> sortedIdx(count);
    int rc = b2gpu_batch_spawn_records(batch, words.data(), offsets.data(), count, sortedIdx.data());
    if (rc != B2GPU_OK) return rc;
    for (int n = 0; n < count; ++n) bodyIndices[order[n]] = sortedIdx[n];
    return B2GPU_OK;
}

int b2gpu_batch_create_bodies(b2gpu_batch* batch, const b2gpu_spawn* spawns, int32_t count, const b2g_fixture_def* fixtures,
                              const float* vertices, int32_t* bodyIndices) {
    return SpawnOrAttach(batch, spawns, nullptr, count, fixtures, vertices, bodyIndices);
}

// b2Body::CreateFixture on bodies that exist already: the same records as a spawn, aimed at an existing body.
int b2gpu_batch_create_fixtures(b2gpu_batch* batch, const b2gpu_attach* attach, int32_t count, const b2g_fixture_def* fixtures,
                                const float* vertices, int32_t* fixtureIndices) {
    if (!batch || count < 0 || (count > 0 && (!attach || !fixtureIndices))) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    std::vector<b2gpu_spawn> spawns((size_t)count);
    std::vector<int32_t> targets((size_t)count);
    for (int i = 0; i < count; ++i) {
        if (attach[i].body < 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "create_fixtures: negative body index");
        memset(&spawns[i], 0, sizeof(b2gpu_spawn));
        spawns[i].world = attach[i].world;
        spawns[i].firstFixture = attach[i].firstFixture;
        spawns[i].fixtureCount = attach[i].fixtureCount;
        spawns[i].body.type = B2G_DYNAMIC_BODY;  // a stand-in: shapes and mass data do not depend on the body
        spawns[i].body.flags = B2G_BD_ACTIVE | B2G_BD_AWAKE;
        spawns[i].body.gravityScale = 1.0f;
        targets[i] = attach[i].body;
    }
    return SpawnOrAttach(batch, spawns.data(), targets.data(), count, fixtures, vertices, fixtureIndices);
}

// The two joint classes whose constructors read the world (b2MouseJoint.cpp:33-52: body B's transform; b2GearJoint.cpp:47-133:
// joint1 / joint2 and the poses of their four bodies): on a running world that state is on the device, so the world's rows
// are read back and the constructor's arithmetic (the same lines as in Flatten above) is evaluated on them.
namespace {
struct LiveWorld {
    int nb = 0, nj = 0;
    std::vector<float> xf, sweep, damp, f0, f1, f2;
    std::vector<int32_t> flags, ji;
};
int ReadLiveWorld(b2gpu_batch* batch, int world, bool joints, LiveWorld& lw) {
    auto rows = [&](int field, auto& dst) -> int {
        int32_t cap = 0, comps = 0;
        int rc = b2gpu_batch_field_shape(batch, field, &cap, &comps);
        if (rc != B2GPU_OK) return rc;
        dst.assign((size_t)cap * comps + 4, 0);
        return cap > 0 ? b2gpu_batch_download_field(batch, field, dst.data(), world, 1) : B2GPU_OK;
    };
    std::vector<int32_t> n;
    int rc = rows(B2GPU_F_BODY_COUNT, n);
    if (rc != B2GPU_OK) return rc;
    lw.nb = n[0];
    if ((rc = rows(B2GPU_F_BODY_XF, lw.xf)) != B2GPU_OK || (rc = rows(B2GPU_F_BODY_SWEEP, lw.sweep)) != B2GPU_OK ||
        (rc = rows(B2GPU_F_BODY_DAMP, lw.damp)) != B2GPU_OK || (rc = rows(B2GPU_F_BODY_FLAGS, lw.flags)) != B2GPU_OK)
        return rc;
    if (!joints) return B2GPU_OK;
    int32_t jointCap = 0;
    if ((rc = b2gpu_batch_field_shape(batch, B2GPU_F_JOINT_I, &jointCap, nullptr)) != B2GPU_OK) return rc;
    if (jointCap == 0) return B2GPU_OK;  // a batch without a joint arena: no joint1 / joint2 to be found
    if ((rc = rows(B2GPU_F_JOINT_COUNT, n)) != B2GPU_OK) return rc;
    lw.nj = n[0];
    if ((rc = rows(B2GPU_F_JOINT_I, lw.ji)) != B2GPU_OK || (rc = rows(B2GPU_F_JOINT_F0, lw.f0)) != B2GPU_OK ||
        (rc = rows(B2GPU_F_JOINT_F1, lw.f1)) != B2GPU_OK || (rc = rows(B2GPU_F_JOINT_F2, lw.f2)) != B2GPU_OK)
        return rc;
    return B2GPU_OK;
}
Xf LiveXf(const LiveWorld& lw, int b) {
    Xf xf;
    xf.p = V2(lw.xf[(size_t)b * 4], lw.xf[(size_t)b * 4 + 1]);
    xf.q.s = lw.xf[(size_t)b * 4 + 2];
    xf.q.c = lw.xf[(size_t)b * 4 + 3];
    return xf;
}
bool LiveBody(const LiveWorld& lw, int b) { return b >= 0 && b < lw.nb && (lw.flags[b] & kBodyDestroyed) == 0; }

int MouseRecord(const LiveWorld& lw, const b2g_joint_def& d, uint32_t* r) {
    auto fb = [](float v) { uint32_t u; memcpy(&u, &v, 4); return u; };
    if (!LiveBody(lw, d.bodyA) || !LiveBody(lw, d.bodyB) || d.bodyA == d.bodyB)
        return b2gpu_host_set_error(B2GPU_ERR_INVALID, "joint references a body that does not exist (or connects a body to itself)");
    const Vec2 target = V2(d.localAnchorA[0], d.localAnchorA[1]);
    const Vec2 lB = MulT(LiveXf(lw, d.bodyB), target);  // m_localAnchorB = b2MulT(m_bodyB->GetTransform(), m_targetA)
    r[0] = (uint32_t)(d.type | ((d.collideConnected ? 1 : 0) << 8));
    r[1] = (uint32_t)(d.bodyA | (d.bodyB << 16));
    r[2] = r[3] = 0;
    r[4] = fb(target.x); r[5] = fb(target.y); r[6] = fb(lB.x); r[7] = fb(lB.y);
    r[8] = fb(d.length); r[9] = fb(d.frequencyHz); r[10] = fb(d.dampingRatio); r[11] = fb(lw.damp[(size_t)d.bodyB * 4 + 3]);
    for (int k = 12; k < 20; ++k) r[k] = 0;
    return B2GPU_OK;
}

int GearRecord(const LiveWorld& lw, const b2g_joint_def& d, uint32_t* r) {
    auto fb = [](float v) { uint32_t u; memcpy(&u, &v, 4); return u; };
    int bodyOf[4];  // A, B, C, D
    float side[2][6] = {{0.0f}};
    int sideType[2];
    float coordinate[2];
    for (int k = 0; k < 2; ++k) {
        const int id = k == 0 ? d.enableLimit : d.enableMotor;  // joint1 / joint2 by joint id
        int row = -1;
        for (int j = 0; j < lw.nj; ++j)
            if (((lw.ji[(size_t)j * 4] >> 16) & 0xffff) == id) row = j;
        if (id < 0 || row < 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "gear joint: joint1 / joint2 must be live joints of the world");
        const int type = lw.ji[(size_t)row * 4] & 0xff;
        if (type != B2G_JOINT_REVOLUTE && type != B2G_JOINT_PRISMATIC)
            return b2gpu_host_set_error(B2GPU_ERR_INVALID, "gear joint: joint1 / joint2 must be revolute or prismatic");
        const int g = lw.ji[(size_t)row * 4 + 1] & 0xffff, m = (lw.ji[(size_t)row * 4 + 1] >> 16) & 0xffff;
        if (!LiveBody(lw, g) || !LiveBody(lw, m)) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "joint references a body that does not exist");
        sideType[k] = type;
        bodyOf[2 + k] = g;  // m_bodyC / m_bodyD = joint->GetBodyA()
        bodyOf[k] = m;      // m_bodyA / m_bodyB = joint->GetBodyB()
        const float* f0 = &lw.f0[(size_t)row * 4];
        const float referenceAngle = lw.f1[(size_t)row * 4];
        if (type == B2G_JOINT_REVOLUTE) {
            side[k][0] = referenceAngle;
            coordinate[k] = lw.sweep[(size_t)m * 4 + 2] - lw.sweep[(size_t)g * 4 + 2] - referenceAngle;
        } else {
            const Vec2 axis = V2(lw.f2[(size_t)row * 4 + 1], lw.f2[(size_t)row * 4 + 2]);  // m_localXAxisA, normalised when it was built
            const Vec2 pG = V2(f0[0], f0[1]), lM = V2(f0[2], f0[3]);
            const Xf ground = LiveXf(lw, g), moving = LiveXf(lw, m);
            const Vec2 pM = MulT(ground.q, Mul(moving.q, lM) + (moving.p - ground.p));
            coordinate[k] = Dot(pM - pG, axis);
            side[k][0] = pG.x; side[k][1] = pG.y; side[k][2] = lM.x; side[k][3] = lM.y; side[k][4] = axis.x; side[k][5] = axis.y;
        }
    }
    if (bodyOf[0] == bodyOf[1]) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "gear joint connects a body to itself");
    const float ratio = d.motorSpeed;
    const float constant = coordinate[0] + ratio * coordinate[1];
    r[0] = (uint32_t)(d.type | ((d.collideConnected ? 1 : 0) << 8) | (sideType[0] << 11) | (sideType[1] << 13));
    r[1] = (uint32_t)(bodyOf[0] | (bodyOf[1] << 16));
    r[2] = 0;
    r[3] = (uint32_t)(bodyOf[2] | (bodyOf[3] << 16));
    for (int k = 0; k < 4; ++k) r[4 + k] = fb(side[0][k]);
    r[8] = fb(side[0][4]); r[9] = fb(side[0][5]); r[10] = fb(ratio); r[11] = fb(constant);
    for (int k = 0; k < 4; ++k) r[12 + k] = fb(side[1][k]);
    r[16] = r[17] = 0;  // m_impulse
    r[18] = fb(side[1][4]); r[19] = fb(side[1][5]);
    return B2GPU_OK;
}
}  // namespace

// b2World::CreateJoint on running worlds: the def's columns, grouped by world, applied by k_create_joints.  A gear or mouse
// joint first sends what the call holds so far (its constructor has to see the joints and edits made before it).
int b2gpu_batch_create_joints(b2gpu_batch* batch, const int32_t* worlds, const b2g_joint_def* defs, int32_t count, int32_t* jointIds) {
    if (!batch || count < 0 || (count > 0 && (!worlds || !defs || !jointIds))) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    if (count == 0) return B2GPU_OK;
    std::vector<int> pending;       // indices into defs, in call order
    std::vector<uint32_t> records;  // 21 words each, parallel to `pending`
    auto send = [&]() -> int {
        const int n = (int)pending.size();
        if (n == 0) return B2GPU_OK;
        std::vector<int> order(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return worlds[pending[a]] < worlds[pending[b]]; });
        std::vector<uint32_t> words((size_t)n * 21);
        for (int i = 0; i < n; ++i) memcpy(&words[(size_t)i * 21], &records[(size_t)order[i] * 21], 21 * sizeof(uint32_t));
        std::vector<int32_t> sorted(n);
        int rc = b2gpu_batch_joint_records(batch, words.data(), n, sorted.data());
        if (rc != B2GPU_OK) return rc;
        for (int i = 0; i < n; ++i) jointIds[pending[order[i]]] = sorted[i];
        pending.clear();
        records.clear();
        return B2GPU_OK;
    };
    for (int i = 0; i < count; ++i) {
        const b2g_joint_def& d = defs[i];
        if (d.type < B2G_JOINT_REVOLUTE || d.type > B2G_JOINT_MOTOR)
            return b2gpu_host_set_error(B2GPU_ERR_UNSUPPORTED, "unknown joint type");
        uint32_t r[21];
        r[0] = (uint32_t)worlds[i];
        if (d.type == B2G_JOINT_GEAR || d.type == B2G_JOINT_MOUSE) {
            int rc = send();
            if (rc != B2GPU_OK) return rc;
            LiveWorld lw;
            if ((rc = ReadLiveWorld(batch, worlds[i], d.type == B2G_JOINT_GEAR, lw)) != B2GPU_OK) return rc;
            rc = d.type == B2G_JOINT_GEAR ? GearRecord(lw, d, r + 1) : MouseRecord(lw, d, r + 1);
            if (rc != B2GPU_OK) return rc;
        } else {
            if (d.bodyA < 0 || d.bodyB < 0 || d.bodyA > 0xffff || d.bodyB > 0xffff || d.bodyA == d.bodyB)
                return b2gpu_host_set_error(B2GPU_ERR_INVALID, "joint references a body that does not exist (or connects a body to itself)");
            PlainJointColumns(d, r + 1);
            for (int k = 17; k < 21; ++k) r[k] = 0;  // accumulated impulses start at zero
        }
        pending.push_back(i);
        records.insert(records.end(), r, r + 21);
    }
    return send();
}

int b2gpu_host_polygon_build(const b2g_fixture_def* fd, const float* vertices, float* out) {
    HostShape sh;
    BuildShape(sh, *fd, vertices);
    for (int i = 0; i < 8; ++i) {
        out[2 * i] = i < sh.count ? sh.verts[i].x : 0.0f;
        out[2 * i + 1] = i < sh.count ? sh.verts[i].y : 0.0f;
        out[16 + 2 * i] = i < sh.count ? sh.normals[i].x : 0.0f;
        out[16 + 2 * i + 1] = i < sh.count ? sh.normals[i].y : 0.0f;
    }
    out[32] = sh.centroid.x;
    out[33] = sh.centroid.y;
    MassData md = ComputeMass(sh, fd->density);
    out[34] = md.mass; out[35] = md.center.x; out[36] = md.center.y; out[37] = md.I;
    return sh.count;
}

int b2gpu_host_fixture_mass_data(const b2g_fixture_def* fd, const float* vertices, int32_t vertexCount, float* out) {
    if (!fd || !out) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    b2g_fixture_def one = *fd;
    one.body = 0;
    int rc = ValidateFixtures(&one, 1, 1, vertices, vertexCount, "fixture");
    if (rc != B2GPU_OK) return rc;
    HostShape sh;
    BuildShape(sh, one, vertices);
    const MassData md = ComputeMass(sh, one.density);
    out[0] = md.mass; out[1] = md.I; out[2] = md.center.x; out[3] = md.center.y;
    return B2GPU_OK;
}

int b2gpu_host_body_mass_data(const b2g_scene* scene, int32_t body, float* out) {
    if (!scene || !out || body < 0 || body >= scene->bodyCount) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    int rc = ValidateFixtures(scene->fixtures, scene->fixtureCount, scene->bodyCount, scene->vertices, scene->vertexCount, "scene");
    if (rc != B2GPU_OK) return rc;
    // b2Body::ResetMassData (b2Body.cpp:283-352), the loop of Flatten above for one body
    const b2g_body_def& bd = scene->bodies[body];
    float mass = 0.0f, I = 0.0f;
    Vec2 localCenter = V2(0.0f, 0.0f);
    if (bd.type == B2G_DYNAMIC_BODY) {
        for (int f = scene->fixtureCount - 1; f >= 0; --f) {  // fixture list is head-inserted: newest first
            const b2g_fixture_def& fd = scene->fixtures[f];
            if (fd.body != body || fd.density == 0.0f) continue;
            HostShape sh;
            BuildShape(sh, fd, scene->vertices);
            MassData md = ComputeMass(sh, fd.density);
            mass += md.mass;
            localCenter = localCenter + md.mass * md.center;
            I += md.I;
        }
        if (mass > 0.0f) localCenter = (1.0f / mass) * localCenter;
        else mass = 1.0f;
        if (I > 0.0f && (bd.flags & B2G_BD_FIXED_ROTATION) == 0) I -= mass * Dot(localCenter, localCenter);
        else I = 0.0f;
    }
    out[0] = mass; out[1] = I; out[2] = localCenter.x; out[3] = localCenter.y;
    return B2GPU_OK;
}

// The branch-free |x| < 120 arithmetic the position solver runs inline (b2g_math.cuh MakeRotWideRange), exposed for the
// bit-parity test against libm; arguments outside that range take SinCos like the solver's fallback does.
int b2gpu_host_sincosf_wide(const float* x, float* s, float* c, int64_t n) {
    if (!x || !s || !c || n < 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    for (int64_t i = 0; i < n; ++i) {
        if (InSinCosWideRange(x[i])) {
            const Rot q = MakeRotWideRange(x[i]);
            s[i] = q.s; c[i] = q.c;
        } else {
            SinCos(x[i], &s[i], &c[i]);
        }
    }
    return B2GPU_OK;
}

int b2gpu_host_sincosf(const float* x, float* s, float* c, int64_t n) {
    if (!x || !s || !c || n < 0) return b2gpu_host_set_error(B2GPU_ERR_INVALID, "bad arguments");
    for (int64_t i = 0; i < n; ++i) SinCos(x[i], &s[i], &c[i]);
    return B2GPU_OK;
}

}  // extern "C"
