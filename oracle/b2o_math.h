/* TEST INFRASTRUCTURE — part of the CPU oracle (plain-C restatement of the reference's
 * b2World::Step path).  Never linked into, imported by or executed from the product path.
 *
 * b2o_math.h: float32 vector / rotation / transform algebra with the reference's exact
 * expression order.  Follows /root/reference/Box2D/Common/b2Math.h (fork) ==
 * Regression/box2d-ref/Box2DRef/Common/b2Math.h (vanilla, cited below as b2Math.h):
 *   b2Dot/b2Cross            b2Math.h:407-430      b2Mul/b2MulT(rot)    b2Math.h:528-537
 *   b2Mul/b2MulT(transform)  b2Math.h:539-555      b2MulT(xf,xf)        b2Math.h:568-575
 *   b2Min/b2Max/b2Clamp      b2Math.h:594-620      Normalize            b2Math.h:101-113
 *   b2Mat22::GetInverse      b2Math.h:211-225      b2Rot::Set           b2Math.h:307-312
 */
#ifndef B2O_MATH_H
#define B2O_MATH_H

#include <float.h>
#include <math.h>
#include <stdint.h>

typedef struct { float x, y; } v2;
typedef struct { float s, c; } rot;
typedef struct { v2 p; rot q; } xform;
typedef struct { v2 lo, hi; } aabb;

#define B2O_PI 3.14159265359f
#define B2O_EPSILON FLT_EPSILON
#define B2O_MAXFLOAT FLT_MAX
#define B2O_LINEAR_SLOP 0.005f
#define B2O_ANGULAR_SLOP (2.0f / 180.0f * B2O_PI)
#define B2O_POLYGON_RADIUS (2.0f * B2O_LINEAR_SLOP)
#define B2O_AABB_EXTENSION 0.1f
#define B2O_AABB_MULTIPLIER 2.0f
#define B2O_VELOCITY_THRESHOLD 1.0f
#define B2O_MAX_LINEAR_CORRECTION 0.2f
#define B2O_MAX_TRANSLATION 2.0f
#define B2O_MAX_TRANSLATION_SQ (B2O_MAX_TRANSLATION * B2O_MAX_TRANSLATION)
#define B2O_MAX_ROTATION (0.5f * B2O_PI)
#define B2O_MAX_ROTATION_SQ (B2O_MAX_ROTATION * B2O_MAX_ROTATION)
#define B2O_BAUMGARTE 0.2f
#define B2O_TIME_TO_SLEEP 0.5f
#define B2O_LINEAR_SLEEP_TOL 0.01f
#define B2O_ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2O_PI)
#define B2O_MAX_POLY 8

static inline v2 V2(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 v2_add(v2 a, v2 b) { return V2(a.x + b.x, a.y + b.y); }
static inline v2 v2_sub(v2 a, v2 b) { return V2(a.x - b.x, a.y - b.y); }
static inline v2 v2_scale(float s, v2 a) { return V2(s * a.x, s * a.y); }
static inline v2 v2_neg(v2 a) { return V2(-a.x, -a.y); }
static inline float v2_dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float v2_cross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 v2_cross_vs(v2 a, float s) { return V2(s * a.y, -s * a.x); }
static inline v2 v2_cross_sv(float s, v2 a) { return V2(-s * a.y, s * a.x); }
static inline float f_min(float a, float b) { return a < b ? a : b; }
static inline float f_max(float a, float b) { return a > b ? a : b; }
static inline float f_clamp(float a, float lo, float hi) { return f_max(lo, f_min(a, hi)); }
static inline float f_abs(float a) { return a > 0.0f ? a : -a; }
static inline v2 v2_min(v2 a, v2 b) { return V2(f_min(a.x, b.x), f_min(a.y, b.y)); }
static inline v2 v2_max(v2 a, v2 b) { return V2(f_max(a.x, b.x), f_max(a.y, b.y)); }
static inline float v2_dist_sq(v2 a, v2 b) { v2 c = v2_sub(a, b); return v2_dot(c, c); }
static inline float v2_len_sq(v2 a) { return a.x * a.x + a.y * a.y; }
static inline float v2_len(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline float v2_normalize(v2* v) {
    float length = v2_len(*v);
    if (length < B2O_EPSILON) return 0.0f;
    float inv = 1.0f / length;
    v->x *= inv;
    v->y *= inv;
    return length;
}
static inline rot rot_set(float angle) { rot q; q.s = sinf(angle); q.c = cosf(angle); return q; }
static inline v2 rot_mul(rot q, v2 v) { return V2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline v2 rot_mulT(rot q, v2 v) { return V2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline rot rot_mulT_rr(rot q, rot r) {
    rot o;
    o.s = q.c * r.s - q.s * r.c;
    o.c = q.c * r.c + q.s * r.s;
    return o;
}
static inline v2 xf_mul(xform T, v2 v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return V2(x, y);
}
static inline v2 xf_mulT(xform T, v2 v) {
    float px = v.x - T.p.x;
    float py = v.y - T.p.y;
    float x = (T.q.c * px + T.q.s * py);
    float y = (-T.q.s * px + T.q.c * py);
    return V2(x, y);
}
static inline xform xf_mulT_xx(xform A, xform B) {
    xform C;
    C.q = rot_mulT_rr(A.q, B.q);
    C.p = rot_mulT(A.q, v2_sub(B.p, A.p));
    return C;
}
static inline int aabb_overlap(aabb a, aabb b) { /* b2TestOverlap, b2Collision.h:278-291 */
    v2 d1 = v2_sub(b.lo, a.hi);
    v2 d2 = v2_sub(a.lo, b.hi);
    if (d1.x > 0.0f || d1.y > 0.0f) return 0;
    if (d2.x > 0.0f || d2.y > 0.0f) return 0;
    return 1;
}
static inline int aabb_contains(aabb a, aabb b) { /* b2AABB::Contains, b2Collision.h:216-224 */
    int r = 1;
    r = r && a.lo.x <= b.lo.x;
    r = r && a.lo.y <= b.lo.y;
    r = r && b.hi.x <= a.hi.x;
    r = r && b.hi.y <= a.hi.y;
    return r;
}

#endif
