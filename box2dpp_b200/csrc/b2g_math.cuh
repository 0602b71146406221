// b2g_math.cuh — float32 algebra of the reference with its exact expression order, for sm_100a.
//
// The build uses --fmad=false -prec-div=true -prec-sqrt=true -ftz=false so that every expression
// below rounds exactly like the reference's x86-64 SSE float path.
// Reference: Box2D/Common/b2Math.h (b2Dot/b2Cross :407-430, b2Mul/b2MulT :528-575,
// b2Min/b2Max/b2Clamp :594-620, Normalize b2Vec.h:204-215, b2Mat22::GetInverse :211-225),
// constants Box2D/Common/b2Settings.h:36-122.
#pragma once

#include <float.h>
#include <stdint.h>

#ifdef __CUDACC__
#define B2G_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define B2G_HD inline
#include <cmath>
#include <cstring>
#endif

namespace b2g {

constexpr float kPi = 3.14159265359f;
constexpr float kEpsilon = FLT_EPSILON;
constexpr float kMaxFloat = FLT_MAX;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = 2.0f / 180.0f * kPi;
constexpr float kPolygonRadius = 2.0f * kLinearSlop;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = 8.0f / 180.0f * kPi;
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = kMaxTranslation * kMaxTranslation;
constexpr float kMaxRotation = 0.5f * kPi;
constexpr float kMaxRotationSquared = kMaxRotation * kMaxRotation;
constexpr float kBaumgarte = 0.2f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = 2.0f / 180.0f * kPi;
constexpr int kMaxPolygonVertices = 8;

struct Vec2 {
    float x, y;
};
struct Rot {
    float s, c;
};
struct Xf {
    Vec2 p;
    Rot q;
};

B2G_HD Vec2 V2(float x, float y) { Vec2 r; r.x = x; r.y = y; return r; }
B2G_HD Vec2 operator+(Vec2 a, Vec2 b) { return V2(a.x + b.x, a.y + b.y); }
B2G_HD Vec2 operator-(Vec2 a, Vec2 b) { return V2(a.x - b.x, a.y - b.y); }
B2G_HD Vec2 operator-(Vec2 a) { return V2(-a.x, -a.y); }
B2G_HD Vec2 operator*(float s, Vec2 a) { return V2(s * a.x, s * a.y); }
B2G_HD float Dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
B2G_HD float Cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
B2G_HD Vec2 Cross(Vec2 a, float s) { return V2(s * a.y, -s * a.x); }
B2G_HD Vec2 Cross(float s, Vec2 a) { return V2(-s * a.y, s * a.x); }
// b2Min / b2Max are ternaries in the reference, NOT fminf/fmaxf (NaN / signed-zero behaviour).
B2G_HD float Min(float a, float b) { return a < b ? a : b; }
B2G_HD float Max(float a, float b) { return a > b ? a : b; }
B2G_HD float Clamp(float a, float lo, float hi) { return Max(lo, Min(a, hi)); }
B2G_HD float Abs(float a) { return a > 0.0f ? a : -a; }
B2G_HD Vec2 Min(Vec2 a, Vec2 b) { return V2(Min(a.x, b.x), Min(a.y, b.y)); }
B2G_HD Vec2 Max(Vec2 a, Vec2 b) { return V2(Max(a.x, b.x), Max(a.y, b.y)); }
B2G_HD float DistanceSquared(Vec2 a, Vec2 b) { Vec2 c = a - b; return Dot(c, c); }
B2G_HD float LengthSquared(Vec2 a) { return a.x * a.x + a.y * a.y; }
B2G_HD float Sqrt(float x) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
B2G_HD float Div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
B2G_HD float Length(Vec2 a) { return Sqrt(a.x * a.x + a.y * a.y); }
B2G_HD float Normalize(Vec2& v) {
    float length = Length(v);
    if (length < kEpsilon) return 0.0f;
    float inv = Div(1.0f, length);
    v.x *= inv;
    v.y *= inv;
    return length;
}
B2G_HD Vec2 Mul(Rot q, Vec2 v) { return V2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
B2G_HD Vec2 MulT(Rot q, Vec2 v) { return V2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
B2G_HD Rot MulT(Rot q, Rot r) {
    Rot o;
    o.s = q.c * r.s - q.s * r.c;
    o.c = q.c * r.c + q.s * r.s;
    return o;
}
B2G_HD Vec2 Mul(Xf T, Vec2 v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return V2(x, y);
}
B2G_HD Vec2 MulT(Xf T, Vec2 v) {
    float px = v.x - T.p.x;
    float py = v.y - T.p.y;
    float x = (T.q.c * px + T.q.s * py);
    float y = (-T.q.s * px + T.q.c * py);
    return V2(x, y);
}
B2G_HD Xf MulT(Xf A, Xf B) {
    Xf C;
    C.q = MulT(A.q, B.q);
    C.p = MulT(A.q, B.p - A.p);
    return C;
}

// ---- sinf / cosf -------------------------------------------------------------------------------
// The reference calls libm sinf/cosf (b2Rot::Set, Box2D/Common/b2Math.h:302-311).  CUDA's sinf/cosf
// round differently, so the device carries its own: the published algorithm of glibc >= 2.28 /
// ARM optimized-routines "sincosf" (double-precision range reduction + degree-7/8 polynomials,
// written here with the FMA contractions of glibc's x86-64 FMA build).  tests/test_sincos.py checks
// it bit-for-bit against this image's libm (exhaustively over all 2^32 inputs in the round-1 probe,
// sampled in CI), which makes body state comparable bit-for-bit with the unmodified reference.
B2G_HD double b2g_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
B2G_HD double b2g_dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
B2G_HD uint32_t AsUint(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

namespace detail {
// 2/pi in 32-bit windows sliding by 8 bits (glibc __inv_pio4)
B2G_HD uint32_t InvPio4(int i) {
    // bits of 2/pi: a2f9836e 4e441529 fc2757d1 f534ddc0 db629599 3c439041 (192 bits, big-endian).
    // Entry i is the 32-bit window that ends at bit 8*(i+1) of that string.
    const uint64_t w0 = 0xa2f9836e4e441529ull, w1 = 0xfc2757d1f534ddc0ull, w2 = 0xdb6295993c439041ull;
    int end = 8 * (i + 1);
    int start = end - 32;
    if (start < 0) return (uint32_t)(w0 >> (64 - end));
    int wi = start >> 6, bo = start & 63;
    uint64_t a = wi == 0 ? w0 : (wi == 1 ? w1 : w2);
    uint64_t b = wi == 0 ? w1 : (wi == 1 ? w2 : 0ull);
    uint64_t v = bo == 0 ? a : ((a << bo) | (b >> (64 - bo)));
    return (uint32_t)(v >> 32);
}
B2G_HD double ReduceLarge(uint32_t xi, int* np) {
    int base = (xi >> 26) & 15;
    int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * InvPio4(base + 0));
    res1 = (uint64_t)xi * InvPio4(base + 4);
    res2 = (uint64_t)xi * InvPio4(base + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    double x = (double)(int64_t)res0;
    *np = (int)n;
    return b2g_dmul(x, 0x1.921FB54442D18p-62);
}
}  // namespace detail

// sin and cos of a float, each bit-identical to glibc's sinf / cosf: every range of the argument.  Kept out of line on
// the device (the large-argument reduction is long and rarely runs); SinCos below handles |y| < pi/4 inline.
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
inline
#endif
void SinCosGeneral(float y, float* sp, float* cp) {
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                 c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    uint32_t top = (AsUint(y) >> 20) & 0x7ff;
    double x = (double)y;
    int n = 0;
    double sgn = 1.0;   // p->sign[n & 3] (with the original sign folded in for the large path)
    bool neg = false;   // use the negated polynomial table
    if (top < ((0x3f490fdbu >> 20) & 0x7ff)) {  // |y| < pi/4
        if (top < ((0x39800000u >> 20) & 0x7ff)) {  // |y| < 2^-12
            *sp = y;
            *cp = 1.0f;
            return;
        }
    } else if (top < ((0x42f00000u >> 20) & 0x7ff)) {  // |y| < 120
        double r = b2g_dmul(x, hpi_inv);
        n = ((int32_t)r + 0x800000) >> 24;
        x = b2g_fma(-(double)n, hpi, x);
        int q = n & 3;
        sgn = (q == 1 || q == 2) ? -1.0 : 1.0;
        neg = (n & 2) != 0;
    } else if (top < 0x7f8) {
        uint32_t xi = AsUint(y);
        int sign = (int)(xi >> 31);
        x = detail::ReduceLarge(xi, &n);
        int q = (n + sign) & 3;
        sgn = (q == 1 || q == 2) ? -1.0 : 1.0;
        neg = ((n + sign) & 2) != 0;
    } else {
        *sp = y - y;  // NaN
        *cp = y - y;
        return;
    }
    double xs = b2g_dmul(x, sgn);
    double x2 = b2g_dmul(x, x);
    // sine polynomial (sinf_poly with even n) and cosine polynomial (odd n)
    double sc0 = neg ? -c0 : c0, sc1 = neg ? -c1 : c1, sc2 = neg ? -c2 : c2, sc3 = neg ? -c3 : c3, sc4 = neg ? -c4 : c4;
    double x3 = b2g_dmul(xs, x2);
    double sA = b2g_fma(x2, s3, s2);
    double x7 = b2g_dmul(x3, x2);
    double sB = b2g_fma(x3, s1, xs);
    float sinPoly = (float)b2g_fma(x7, sA, sB);
    double x4 = b2g_dmul(x2, x2);
    double cA = b2g_fma(x2, sc4, sc3);
    double cB = b2g_fma(x2, sc1, sc0);
    double x6 = b2g_dmul(x4, x2);
    double cC = b2g_fma(x4, sc2, cB);
    float cosPoly = (float)b2g_fma(x6, cA, cC);
    // sinf uses poly(n), cosf uses poly(n ^ 1)
    if ((n & 1) == 0) {
        *sp = sinPoly;
        *cp = cosPoly;
    } else {
        *sp = cosPoly;
        *cp = sinPoly;
    }
}

// sinf / cosf as above.  |y| < pi/4 (every resting or slowly turning body) runs inline and branch-free: it is the
// general path with n = 0, sign = +1 and the plain polynomial table (x * 1.0 is exact, so `xs` is `x`), and the
// |y| < 2^-12 shortcut (sin = y, cos = 1) becomes a select.
B2G_HD void SinCos(float y, float* sp, float* cp) {
    const uint32_t top = (AsUint(y) >> 20) & 0x7ff;
    if (top < ((0x3f490fdbu >> 20) & 0x7ff)) {
        const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                     c4 = 0x1.99343027bf8c3p-16;
        const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
        const double x = (double)y;
        const double x2 = b2g_dmul(x, x);
        const double x3 = b2g_dmul(x, x2);
        const double sA = b2g_fma(x2, s3, s2);
        const double x7 = b2g_dmul(x3, x2);
        const double sB = b2g_fma(x3, s1, x);
        const float sinPoly = (float)b2g_fma(x7, sA, sB);
        const double x4 = b2g_dmul(x2, x2);
        const double cA = b2g_fma(x2, c4, c3);
        const double cB = b2g_fma(x2, c1, c0);
        const double x6 = b2g_dmul(x4, x2);
        const double cC = b2g_fma(x4, c2, cB);
        const float cosPoly = (float)b2g_fma(x6, cA, cC);
        const bool tiny = top < ((0x39800000u >> 20) & 0x7ff);  // |y| < 2^-12
        *sp = tiny ? y : sinPoly;
        *cp = tiny ? 1.0f : cosPoly;
        return;
    }
    SinCosGeneral(y, sp, cp);
}

// The |y| < pi/4 arithmetic of SinCos without the range test (InSinCosFastRange is the caller's business): the sequential
// solver evaluates it unconditionally and redoes the contact through SinCos when an angle was out of range.
B2G_HD bool InSinCosFastRange(float y) { return ((AsUint(y) >> 20) & 0x7ff) < ((0x3f490fdbu >> 20) & 0x7ff); }
B2G_HD Rot MakeRotFastRange(float y) {
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                 c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double x = (double)y;
    const double x2 = b2g_dmul(x, x);
    const double x3 = b2g_dmul(x, x2);
    const double sA = b2g_fma(x2, s3, s2);
    const double x7 = b2g_dmul(x3, x2);
    const double sB = b2g_fma(x3, s1, x);
    const float sinPoly = (float)b2g_fma(x7, sA, sB);
    const double x4 = b2g_dmul(x2, x2);
    const double cA = b2g_fma(x2, c4, c3);
    const double cB = b2g_fma(x2, c1, c0);
    const double x6 = b2g_dmul(x4, x2);
    const double cC = b2g_fma(x4, c2, cB);
    const float cosPoly = (float)b2g_fma(x6, cA, cC);
    const bool tiny = ((AsUint(y) >> 20) & 0x7ff) < ((0x39800000u >> 20) & 0x7ff);  // |y| < 2^-12
    Rot q;
    q.s = tiny ? y : sinPoly;
    q.c = tiny ? 1.0f : cosPoly;
    return q;
}

// The same for |y| < 120 (the range glibc reduces with one multiply and one fused multiply-add, SinCosGeneral's second
// branch), still branch-free: quadrant n = round(y * 2/pi), reduced argument x - n * pi/2, the two polynomials on it, then
// sign and swap by quadrant.  Negating every coefficient of the cosine polynomial (glibc's second table) negates its value
// exactly, and the sine polynomial is odd in its argument, so the signs are applied to the results.  Bodies that have
// turned further than pi/4 (anything that rolls or tumbles) stay on the sequential solver's straight-line path with it.
B2G_HD bool InSinCosWideRange(float y) { return ((AsUint(y) >> 20) & 0x7ff) < ((0x42f00000u >> 20) & 0x7ff); }
B2G_HD Rot MakeRotWideRange(float y) {
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                 c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const uint32_t top = (AsUint(y) >> 20) & 0x7ff;
    const bool reduce = top >= ((0x3f490fdbu >> 20) & 0x7ff);  // |y| >= pi/4
    const double x0 = (double)y;
    const double r = b2g_dmul(x0, hpi_inv);
    const int nr = ((int32_t)r + 0x800000) >> 24;
    const int n = reduce ? nr : 0;
    const double xr = b2g_fma(-(double)nr, hpi, x0);
    const double x = reduce ? xr : x0;
    const double x2 = b2g_dmul(x, x);
    const double x3 = b2g_dmul(x, x2);
    const double sA = b2g_fma(x2, s3, s2);
    const double x7 = b2g_dmul(x3, x2);
    const double sB = b2g_fma(x3, s1, x);
    float sinPoly = (float)b2g_fma(x7, sA, sB);
    const double x4 = b2g_dmul(x2, x2);
    const double cA = b2g_fma(x2, c4, c3);
    const double cB = b2g_fma(x2, c1, c0);
    const double x6 = b2g_dmul(x4, x2);
    const double cC = b2g_fma(x4, c2, cB);
    float cosPoly = (float)b2g_fma(x6, cA, cC);
    const int q = n & 3;
    sinPoly = (q == 1 || q == 2) ? -sinPoly : sinPoly;  // sign[n & 3] multiplies the odd polynomial's argument
    cosPoly = (n & 2) ? -cosPoly : cosPoly;             // the negated coefficient table
    const bool tiny = top < ((0x39800000u >> 20) & 0x7ff);  // |y| < 2^-12
    Rot o;
    o.s = tiny ? y : ((n & 1) ? cosPoly : sinPoly);
    o.c = tiny ? 1.0f : ((n & 1) ? sinPoly : cosPoly);
    return o;
}

B2G_HD Rot MakeRot(float angle) {
    Rot q;
    SinCos(angle, &q.s, &q.c);
    return q;
}

}  // namespace b2g
