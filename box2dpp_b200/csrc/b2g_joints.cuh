// b2g_joints.cuh — joints in the island solver: revolute, prismatic, distance, wheel, weld, rope, friction, motor, pulley, mouse and gear joints (all 11 of b2JointType).
//
// Reference (vanilla twin Regression/box2d-ref/Box2DRef/Dynamics/Joints; the fork's files are
// Box2D/Dynamics/Joints/b2RevoluteJoint.cpp, b2PrismaticJoint.cpp, b2DistanceJoint.cpp, same line structure):
//   InitVelocityConstraints / SolveVelocityConstraints / SolvePositionConstraints of each joint class,
//   b2Mat33::Solve33 / Solve22 (Common/b2Math.cpp:22-52), b2Mat22::Solve (Common/b2Math.h:242-254).
// Hooks in the island solver: b2Island.cpp:252-255 (init after the contact warm start), :263-266 (before the contacts in
// every velocity iteration), :316-320 (after the contacts in every position iteration).
//
// One world per thread, like the rest of the generic solver kernel (k_solve): positions and velocities are the body
// arrays (b2Position = bodySweep.c/a, b2Velocity = bodyVel), joint temporaries live in a per-world scratch column.
// Textually included inside namespace b2g (from b2g_kernels.cuh).
#pragma once

constexpr int kJointTmp = 5;  // float4 of solver temporaries per joint
enum : int { kLimitInactive = 0, kLimitAtLower = 1, kLimitAtUpper = 2, kLimitEqual = 3 };  // b2LimitState
enum : int { kJointCollide = 1 << 8, kJointEnableLimit = 1 << 9, kJointEnableMotor = 1 << 10 };

struct Vec3 { float x, y, z; };
__device__ __forceinline__ Vec3 V3(float x, float y, float z) { Vec3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ float Dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 Cross3(Vec3 a, Vec3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
struct Mat33 { Vec3 ex, ey, ez; };
// b2Mat33::Solve33 — Common/b2Math.cpp:22-35
__device__ __forceinline__ Vec3 Solve33(const Mat33& M, Vec3 b) {
    float det = Dot3(M.ex, Cross3(M.ey, M.ez));
    if (det != 0.0f) det = Div(1.0f, det);
    Vec3 x;
    x.x = det * Dot3(b, Cross3(M.ey, M.ez));
    x.y = det * Dot3(M.ex, Cross3(b, M.ez));
    x.z = det * Dot3(M.ex, Cross3(M.ey, b));
    return x;
}
// b2Mat33::Solve22 / b2Mat22::Solve — Common/b2Math.cpp:39-52, b2Math.h:242-254 (a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y)
__device__ __forceinline__ Vec2 Solve22(float a11, float a12, float a21, float a22, Vec2 b) {
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = Div(1.0f, det);
    Vec2 x;
    x.x = det * (a22 * b.x - a12 * b.y);
    x.y = det * (a11 * b.y - a21 * b.x);
    return x;
}

// b2Mat33::GetInverse22 / GetSymInverse33 — Common/b2Math.cpp:55-95; b2Mul(b2Mat33, b2Vec3), b2Mul22 — b2Math.h:531-540
__device__ __forceinline__ Mat33 GetInverse22(const Mat33& K) {
    const float a = K.ex.x, b = K.ey.x, c = K.ex.y, d = K.ey.y;
    float det = a * d - b * c;
    if (det != 0.0f) det = Div(1.0f, det);
    Mat33 M;
    M.ex = V3(det * d, -det * c, 0.0f);
    M.ey = V3(-det * b, det * a, 0.0f);
    M.ez = V3(0.0f, 0.0f, 0.0f);
    return M;
}
__device__ __forceinline__ Mat33 GetSymInverse33(const Mat33& K) {
    float det = Dot3(K.ex, Cross3(K.ey, K.ez));
    if (det != 0.0f) det = Div(1.0f, det);
    const float a11 = K.ex.x, a12 = K.ey.x, a13 = K.ez.x;
    const float a22 = K.ey.y, a23 = K.ez.y;
    const float a33 = K.ez.z;
    Mat33 M;
    M.ex.x = det * (a22 * a33 - a23 * a23);
    M.ex.y = det * (a13 * a23 - a12 * a33);
    M.ex.z = det * (a12 * a23 - a13 * a22);
    M.ey.x = M.ex.y;
    M.ey.y = det * (a11 * a33 - a13 * a13);
    M.ey.z = det * (a13 * a12 - a11 * a23);
    M.ez.x = M.ex.z;
    M.ez.y = M.ey.z;
    M.ez.z = det * (a11 * a22 - a12 * a12);
    return M;
}
__device__ __forceinline__ Vec3 Mul33(const Mat33& A, Vec3 v) {
    return V3(v.x * A.ex.x + v.y * A.ey.x + v.z * A.ez.x, v.x * A.ex.y + v.y * A.ey.y + v.z * A.ez.y,
              v.x * A.ex.z + v.y * A.ey.z + v.z * A.ez.z);
}
__device__ __forceinline__ Vec2 Mul22(const Mat33& A, Vec2 v) {
    return V2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}

// Where a joint finds its bodies' b2Velocity / b2Position during the solve.  Generic solver (k_solve): the body arrays in
// global memory are the working copies.  Tile solver (b2g_solve.cuh): the tile's table in shared memory, row b at
// table + b * 384, components 128 bytes apart, this lane's column already added in (0 = not the table).
struct JointCols {
    Col<int4> jI;
    Col<float4> jF0, jF1, jF2, jImp, jTmp;
    Col<float4> bodySweep, bodyVel, bodyMass, bodyDamp;  // bodyDamp.w = m_mass (b2Body::GetMass, read by the mouse joint)
    uint32_t velTable, posTable;
    static __device__ __forceinline__ float4 LoadRow(uint32_t a) {
        float4 v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v.x) : "r"(a));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v.y) : "r"(a + 128));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v.z) : "r"(a + 256));
        v.w = 0.0f;
        return v;
    }
    static __device__ __forceinline__ void StoreRow(uint32_t a, float4 v) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v.x));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 128), "f"(v.y));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 256), "f"(v.z));
    }
    __device__ __forceinline__ float4 vel(int b) const { return velTable ? LoadRow(velTable + (uint32_t)b * 384u) : bodyVel[b]; }
    __device__ __forceinline__ void setVel(int b, float4 v) const {
        if (velTable) StoreRow(velTable + (uint32_t)b * 384u, v);
        else bodyVel[b] = v;
    }
    // m_sweep.c.xy, m_sweep.a (.w, the sleep time, is not the solver's: it stays where it is on the table path)
    __device__ __forceinline__ float4 sweep(int b) const { return posTable ? LoadRow(posTable + (uint32_t)b * 384u) : bodySweep[b]; }
    __device__ __forceinline__ void setSweep(int b, float4 v) const {
        if (posTable) StoreRow(posTable + (uint32_t)b * 384u, v);
        else bodySweep[b] = v;
    }
};
__device__ __forceinline__ JointCols MakeJointCols(const Arena& A, int tile, int lane, uint32_t velTable = 0, uint32_t posTable = 0) {
    JointCols c{col(A.jointI, tile, A.NJ, lane), col(A.jointF0, tile, A.NJ, lane), col(A.jointF1, tile, A.NJ, lane),
                col(A.jointF2, tile, A.NJ, lane), col(A.jointImp, tile, A.NJ, lane), col(A.jointTmp, tile, A.NJ * kJointTmp, lane),
                col(A.bodySweep, tile, A.NB, lane), col(A.bodyVel, tile, A.NB, lane), col(A.bodyMass, tile, A.NB, lane),
                col(A.bodyDamp, tile, A.NB, lane), velTable, posTable};
    return c;
}


// ---- gear joint (b2GearJoint.cpp): four bodies.  Record: jI.x bits 11-12 / 13-14 = type of joint1 / joint2, jI.y = A | B << 16,
// jI.w = C | D << 16 (A, C = joint1's bodyB, bodyA; B, D = joint2's).  Per side six constants: revolute (referenceAngle, -),
// prismatic (localAnchor on C / D .xy, localAnchor on A / B .xy, localAxis .xy).  Side 1 = jF0.xyzw + jF1.xy, side 2 =
// jF2.xyzw + jImp.zw; jF1.z = ratio, jF1.w = m_constant; jImp.x = m_impulse.
struct GearJacobian {
    Vec2 JvAC, JvBD;
    float JwA, JwB, JwC, JwD, mass;
};
// ---- InitVelocityConstraints ---------------------------------------------------------------------------------
__device__ __noinline__ void JointInitVelocity(const JointCols* cp, int j, float h, float dtRatio, bool warmStarting) {
    const JointCols& c = *cp;
    int4 ji = c.jI[j];
    const int type = ji.x & 0xff;
    const int bA = ji.y & 0xffff, bB = (ji.y >> 16) & 0xffff;
    const float4 f0 = c.jF0[j], f1 = c.jF1[j], f2 = c.jF2[j];
    const float4 msA = c.bodyMass[bA], msB = c.bodyMass[bB];
    const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
    const float4 swA = c.sweep(bA), swB = c.sweep(bB);
    const Vec2 cA = V2(swA.x, swA.y), cB = V2(swB.x, swB.y);
    const float aA = swA.z, aB = swB.z;
    float4 velA = c.vel(bA), velB = c.vel(bB);
    Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
    float wA = velA.z, wB = velB.z;
    const Rot qA = MakeRot(aA), qB = MakeRot(aB);
    const Vec2 rA = Mul(qA, V2(f0.x, f0.y) - V2(msA.z, msA.w));
    const Vec2 rB = Mul(qB, V2(f0.z, f0.w) - V2(msB.z, msB.w));
    float4 imp = c.jImp[j];
    const size_t t = (size_t)j * kJointTmp;

    if (type == B2G_JOINT_GEAR) {  // b2GearJoint.cpp:135-248
        const int bC = ji.w & 0xffff, bD = (ji.w >> 16) & 0xffff;
        const int typeA = (ji.x >> 11) & 3, typeB = (ji.x >> 13) & 3;
        const float4 msC = c.bodyMass[bC], msD = c.bodyMass[bD];
        const float mC = msC.x, iC = msC.y, mD = msD.x, iD = msD.y;
        const float aC = c.sweep(bC).z, aD = c.sweep(bD).z;
        float4 velC = c.vel(bC), velD = c.vel(bD);
        Vec2 vC = V2(velC.x, velC.y), vD = V2(velD.x, velD.y);
        float wC = velC.z, wD = velD.z;
        const Rot qC = MakeRot(aC), qD = MakeRot(aD);
        const float ratio = f1.z;
        GearJacobian g;
        g.mass = 0.0f;
        if (typeA == B2G_JOINT_REVOLUTE) {
            g.JvAC = V2(0.0f, 0.0f);
            g.JwA = 1.0f;
            g.JwC = 1.0f;
            g.mass += iA + iC;
        } else {
            const Vec2 u = Mul(qC, V2(f1.x, f1.y));
            const Vec2 rC = Mul(qC, V2(f0.x, f0.y) - V2(msC.z, msC.w));
            const Vec2 gA = Mul(qA, V2(f0.z, f0.w) - V2(msA.z, msA.w));
            g.JvAC = u;
            g.JwC = Cross(rC, u);
            g.JwA = Cross(gA, u);
            g.mass += mC + mA + iC * g.JwC * g.JwC + iA * g.JwA * g.JwA;
        }
        if (typeB == B2G_JOINT_REVOLUTE) {
            g.JvBD = V2(0.0f, 0.0f);
            g.JwB = ratio;
            g.JwD = ratio;
            g.mass += ratio * ratio * (iB + iD);
        } else {
            const Vec2 u = Mul(qD, V2(imp.z, imp.w));
            const Vec2 rD = Mul(qD, V2(f2.x, f2.y) - V2(msD.z, msD.w));
            const Vec2 gB = Mul(qB, V2(f2.z, f2.w) - V2(msB.z, msB.w));
            g.JvBD = ratio * u;
            g.JwD = ratio * Cross(rD, u);
            g.JwB = ratio * Cross(gB, u);
            g.mass += ratio * ratio * (mD + mB) + iD * g.JwD * g.JwD + iB * g.JwB * g.JwB;
        }
        g.mass = g.mass > 0.0f ? Div(1.0f, g.mass) : 0.0f;
        if (warmStarting) {  // (no dtRatio scaling in the reference)
            vA = vA + (mA * imp.x) * g.JvAC;
            wA += iA * imp.x * g.JwA;
            vB = vB + (mB * imp.x) * g.JvBD;
            wB += iB * imp.x * g.JwB;
            vC = vC - (mC * imp.x) * g.JvAC;
            wC -= iC * imp.x * g.JwC;
            vD = vD - (mD * imp.x) * g.JvBD;
            wD -= iD * imp.x * g.JwD;
        } else {
            imp.x = 0.0f;
        }
        c.jTmp[t + 0] = make_float4(g.JvAC.x, g.JvAC.y, g.JvBD.x, g.JvBD.y);
        c.jTmp[t + 1] = make_float4(g.JwA, g.JwB, g.JwC, g.JwD);
        c.jTmp[t + 2] = make_float4(g.mass, 0.0f, 0.0f, 0.0f);
        c.jImp[j] = imp;
        // stores in the reference's order A, B, C, D (C and D may be the same body: the later store wins)
        c.setVel(bA, make_float4(vA.x, vA.y, wA, 0.0f));
        c.setVel(bB, make_float4(vB.x, vB.y, wB, 0.0f));
        c.setVel(bC, make_float4(vC.x, vC.y, wC, 0.0f));
        c.setVel(bD, make_float4(vD.x, vD.y, wD, 0.0f));
        return;
    }
    if (type == B2G_JOINT_MOUSE) {  // b2MouseJoint.cpp:96-160; f0 = target.xy, localAnchorB.xy; f1 = maxForce, frequencyHz, dampingRatio, mass of B
        const float mass = c.bodyDamp[bB].w;  // m_bodyB->GetMass() as it is now (b2Body::SetMassData / ResetMassData may have changed it)
        const float omega = 2.0f * kPi * f1.y;
        const float d = 2.0f * mass * f1.z * omega;
        const float k = mass * (omega * omega);
        float gamma = h * (d + h * k);
        if (gamma != 0.0f) gamma = Div(1.0f, gamma);
        const float beta = h * k * gamma;
        const float kxx = mB + iB * rB.y * rB.y + gamma;  // K.ex.x
        const float kxy = -iB * rB.x * rB.y;              // K.ex.y = K.ey.x
        const float kyy = mB + iB * rB.x * rB.x + gamma;  // K.ey.y
        float det = kxx * kyy - kxy * kxy;                // b2Mat22::GetInverse (b2Math.h:226-238)
        if (det != 0.0f) det = Div(1.0f, det);
        const float mxx = det * kyy, myx = -det * kxy, mxy = -det * kxy, myy = det * kxx;  // ex.x, ey.x, ex.y, ey.y
        Vec2 C = cB + rB - V2(f0.x, f0.y);
        C.x *= beta;
        C.y *= beta;
        wB *= 0.98f;
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio;
            const Vec2 P = V2(imp.x, imp.y);
            vB = vB + mB * P;
            wB += iB * Cross(rB, P);
        } else {
            imp.x = 0.0f; imp.y = 0.0f;
        }
        c.jTmp[t + 0] = make_float4(rB.x, rB.y, C.x, C.y);
        c.jTmp[t + 1] = make_float4(mxx, myx, mxy, myy);
        c.jTmp[t + 2] = make_float4(gamma, 0.0f, 0.0f, 0.0f);
    } else if (type == B2G_JOINT_REVOLUTE) {  // b2RevoluteJoint.cpp:66-176
        const bool fixedRotation = (iA + iB == 0.0f);
        Mat33 M;
        M.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        M.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        M.ez.x = -rA.y * iA - rB.y * iB;
        M.ex.y = M.ey.x;
        M.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        M.ez.y = rA.x * iA + rB.x * iB;
        M.ex.z = M.ez.x;
        M.ey.z = M.ez.y;
        M.ez.z = iA + iB;
        float motorMass = iA + iB;
        if (motorMass > 0.0f) motorMass = Div(1.0f, motorMass);
        if ((ji.x & kJointEnableMotor) == 0 || fixedRotation) imp.w = 0.0f;
        int limitState = ji.z;
        if ((ji.x & kJointEnableLimit) && !fixedRotation) {
            const float jointAngle = aB - aA - f1.x;
            if (Abs(f1.z - f1.y) < 2.0f * kAngularSlop) {
                limitState = kLimitEqual;
            } else if (jointAngle <= f1.y) {
                if (limitState != kLimitAtLower) imp.z = 0.0f;
                limitState = kLimitAtLower;
            } else if (jointAngle >= f1.z) {
                if (limitState != kLimitAtUpper) imp.z = 0.0f;
                limitState = kLimitAtUpper;
            } else {
                limitState = kLimitInactive;
                imp.z = 0.0f;
            }
        } else {
            limitState = kLimitInactive;
        }
        if (limitState != ji.z) {
            ji.z = limitState;
            c.jI[j] = ji;
        }
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
            imp.w *= dtRatio;
            const Vec2 P = V2(imp.x, imp.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(rA, P) + imp.w + imp.z);
            vB = vB + mB * P;
            wB += iB * (Cross(rB, P) + imp.w + imp.z);
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        c.jTmp[t + 1] = make_float4(M.ex.x, M.ex.y, M.ex.z, motorMass);
        c.jTmp[t + 2] = make_float4(M.ey.x, M.ey.y, M.ey.z, 0.0f);
        c.jTmp[t + 3] = make_float4(M.ez.x, M.ez.y, M.ez.z, 0.0f);
    } else if (type == B2G_JOINT_DISTANCE) {  // b2DistanceJoint.cpp:64-152
        Vec2 u = cB + rB - cA - rA;
        const float length = Length(u);
        if (length > kLinearSlop) u = Div(1.0f, length) * u;
        else u = V2(0.0f, 0.0f);
        const float crAu = Cross(rA, u), crBu = Cross(rB, u);
        float invMass = mA + iA * crAu * crAu + mB + iB * crBu * crBu;
        float mass = invMass != 0.0f ? Div(1.0f, invMass) : 0.0f;
        float gamma = 0.0f, bias = 0.0f;
        if (f1.y > 0.0f) {
            const float C = length - f1.x;
            const float omega = 2.0f * kPi * f1.y;
            const float d = 2.0f * mass * f1.z * omega;
            const float k = mass * omega * omega;
            gamma = h * (d + h * k);
            gamma = gamma != 0.0f ? Div(1.0f, gamma) : 0.0f;
            bias = C * h * k * gamma;
            invMass += gamma;
            mass = invMass != 0.0f ? Div(1.0f, invMass) : 0.0f;
        }
        if (warmStarting) {
            imp.x *= dtRatio;
            const Vec2 P = imp.x * u;
            vA = vA - mA * P;
            wA -= iA * Cross(rA, P);
            vB = vB + mB * P;
            wB += iB * Cross(rB, P);
        } else {
            imp.x = 0.0f;
        }
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        c.jTmp[t + 1] = make_float4(u.x, u.y, mass, gamma);
        c.jTmp[t + 2] = make_float4(bias, 0.0f, 0.0f, 0.0f);
    } else if (type == B2G_JOINT_PULLEY) {  // b2PulleyJoint.cpp:82-164; f1 = groundAnchorA.xy, groundAnchorB.xy; f2 = ratio, constant
        Vec2 uA = cA + rA - V2(f1.x, f1.y);
        Vec2 uB = cB + rB - V2(f1.z, f1.w);
        const float lengthA = Length(uA), lengthB = Length(uB);
        if (lengthA > 10.0f * kLinearSlop) uA = Div(1.0f, lengthA) * uA; else uA = V2(0.0f, 0.0f);
        if (lengthB > 10.0f * kLinearSlop) uB = Div(1.0f, lengthB) * uB; else uB = V2(0.0f, 0.0f);
        const float ruA = Cross(rA, uA), ruB = Cross(rB, uB);
        const float pmA = mA + iA * ruA * ruA;
        const float pmB = mB + iB * ruB * ruB;
        float mass = pmA + f2.x * f2.x * pmB;
        if (mass > 0.0f) mass = Div(1.0f, mass);
        if (warmStarting) {
            imp.x *= dtRatio;
            const Vec2 PA = -(imp.x) * uA;
            const Vec2 PB = (-f2.x * imp.x) * uB;
            vA = vA + mA * PA;
            wA += iA * Cross(rA, PA);
            vB = vB + mB * PB;
            wB += iB * Cross(rB, PB);
        } else {
            imp.x = 0.0f;
        }
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        c.jTmp[t + 1] = make_float4(uA.x, uA.y, uB.x, uB.y);
        c.jTmp[t + 2] = make_float4(mass, 0.0f, 0.0f, 0.0f);
    } else if (type == B2G_JOINT_MOTOR) {  // b2MotorJoint.cpp:61-133; f0.xy = linearOffset, f1 = maxForce, maxTorque, angularOffset, correctionFactor
        const Vec2 mrA = Mul(qA, -V2(msA.z, msA.w)), mrB = Mul(qB, -V2(msB.z, msB.w));
        const float k11 = mA + mB + iA * mrA.y * mrA.y + iB * mrB.y * mrB.y;
        const float k12 = -iA * mrA.x * mrA.y - iB * mrB.x * mrB.y;
        const float k22 = mA + mB + iA * mrA.x * mrA.x + iB * mrB.x * mrB.x;
        float det = k11 * k22 - k12 * k12;
        if (det != 0.0f) det = Div(1.0f, det);
        const float lmxx = det * k22, lmyx = -det * k12, lmxy = -det * k12, lmyy = det * k11;
        float angularMass = iA + iB;
        if (angularMass > 0.0f) angularMass = Div(1.0f, angularMass);
        const Vec2 linearError = cB + mrB - cA - mrA - Mul(qA, V2(f0.x, f0.y));
        const float angularError = aB - aA - f1.z;
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio;
            imp.z *= dtRatio;
            const Vec2 P = V2(imp.x, imp.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(mrA, P) + imp.z);
            vB = vB + mB * P;
            wB += iB * (Cross(mrB, P) + imp.z);
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(mrA.x, mrA.y, mrB.x, mrB.y);
        c.jTmp[t + 1] = make_float4(lmxx, lmyx, lmxy, lmyy);
        c.jTmp[t + 2] = make_float4(angularMass, linearError.x, linearError.y, angularError);
    } else if (type == B2G_JOINT_ROPE) {  // b2RopeJoint.cpp:59-133; f1.x = maxLength
        Vec2 u = cB + rB - cA - rA;
        const float length = Length(u);
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        if (length > kLinearSlop) {
            u = Div(1.0f, length) * u;
        } else {  // degenerate rope: no constraint this step (the reference returns before touching the velocities)
            c.jTmp[t + 1] = make_float4(0.0f, 0.0f, 0.0f, length);
            imp.x = 0.0f;
            c.jImp[j] = imp;
            return;
        }
        const float crA = Cross(rA, u), crB = Cross(rB, u);
        const float invMass = mA + iA * crA * crA + mB + iB * crB * crB;
        const float mass = invMass != 0.0f ? Div(1.0f, invMass) : 0.0f;
        if (warmStarting) {
            imp.x *= dtRatio;
            const Vec2 P = imp.x * u;
            vA = vA - mA * P;
            wA -= iA * Cross(rA, P);
            vB = vB + mB * P;
            wB += iB * Cross(rB, P);
        } else {
            imp.x = 0.0f;
        }
        c.jTmp[t + 1] = make_float4(u.x, u.y, mass, length);
    } else if (type == B2G_JOINT_FRICTION) {  // b2FrictionJoint.cpp:63-128; f1 = maxForce, maxTorque; jImp = linear.xy, angular
        const float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        const float k12 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        const float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        // b2Mat22::GetInverse (b2Math.h:226-238) with a = ex.x, b = ey.x, c = ex.y, d = ey.y
        float det = k11 * k22 - k12 * k12;
        if (det != 0.0f) det = Div(1.0f, det);
        const float lmxx = det * k22, lmyx = -det * k12, lmxy = -det * k12, lmyy = det * k11;  // ex.x, ey.x, ex.y, ey.y
        float angularMass = iA + iB;
        if (angularMass > 0.0f) angularMass = Div(1.0f, angularMass);
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio;
            imp.z *= dtRatio;
            const Vec2 P = V2(imp.x, imp.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(rA, P) + imp.z);
            vB = vB + mB * P;
            wB += iB * (Cross(rB, P) + imp.z);
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        c.jTmp[t + 1] = make_float4(lmxx, lmyx, lmxy, lmyy);
        c.jTmp[t + 2] = make_float4(angularMass, 0.0f, 0.0f, 0.0f);
    } else if (type == B2G_JOINT_WHEEL) {  // b2WheelJoint.cpp:79-199
        // jImp: m_impulse, m_springImpulse, -, m_motorImpulse.  f1: frequencyHz, dampingRatio, -, maxMotorTorque
        const Vec2 d = cB + rB - cA - rA;
        const Vec2 localX = V2(f2.y, f2.z), localY = Cross(1.0f, localX);
        const Vec2 ay = Mul(qA, localY);
        const float sAy = Cross(d + rA, ay), sBy = Cross(rB, ay);
        float mass = mA + mB + iA * sAy * sAy + iB * sBy * sBy;
        if (mass > 0.0f) mass = Div(1.0f, mass);
        float springMass = 0.0f, bias = 0.0f, gamma = 0.0f;
        // m_ax / m_sAx / m_sBx keep their constructor values (zero) while the spring is off
        Vec2 ax = V2(0.0f, 0.0f);
        float sAx = 0.0f, sBx = 0.0f;
        if (f1.x > 0.0f) {
            ax = Mul(qA, localX);
            sAx = Cross(d + rA, ax);
            sBx = Cross(rB, ax);
            const float invMass = mA + mB + iA * sAx * sAx + iB * sBx * sBx;
            if (invMass > 0.0f) {
                springMass = Div(1.0f, invMass);
                const float C = Dot(d, ax);
                const float omega = 2.0f * kPi * f1.x;
                const float dd = 2.0f * springMass * f1.y * omega;
                const float k = springMass * omega * omega;
                gamma = h * (dd + h * k);
                if (gamma > 0.0f) gamma = Div(1.0f, gamma);
                bias = C * h * k * gamma;
                springMass = invMass + gamma;
                if (springMass > 0.0f) springMass = Div(1.0f, springMass);
            }
        } else {
            imp.y = 0.0f;
        }
        float motorMass = 0.0f;
        if (ji.x & kJointEnableMotor) {
            motorMass = iA + iB;
            if (motorMass > 0.0f) motorMass = Div(1.0f, motorMass);
        } else {
            imp.w = 0.0f;
        }
        if (warmStarting) {
            imp.x *= dtRatio;
            imp.y *= dtRatio;
            imp.w *= dtRatio;
            const Vec2 P = imp.x * ay + imp.y * ax;
            const float LA = imp.x * sAy + imp.y * sAx + imp.w;
            const float LB = imp.x * sBy + imp.y * sBx + imp.w;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(ax.x, ax.y, ay.x, ay.y);
        c.jTmp[t + 1] = make_float4(sAx, sBx, sAy, sBy);
        c.jTmp[t + 2] = make_float4(mass, springMass, motorMass, gamma);
        c.jTmp[t + 3] = make_float4(bias, 0.0f, 0.0f, 0.0f);
    } else if (type == B2G_JOINT_WELD) {  // b2WeldJoint.cpp:66-154
        // f1: referenceAngle, frequencyHz, dampingRatio
        Mat33 K;
        K.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        K.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        K.ez.x = -rA.y * iA - rB.y * iB;
        K.ex.y = K.ey.x;
        K.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        K.ez.y = rA.x * iA + rB.x * iB;
        K.ex.z = K.ez.x;
        K.ey.z = K.ez.y;
        K.ez.z = iA + iB;
        Mat33 M;
        float gamma = 0.0f, bias = 0.0f;
        if (f1.y > 0.0f) {
            M = GetInverse22(K);
            float invM = iA + iB;
            const float m = invM > 0.0f ? Div(1.0f, invM) : 0.0f;
            const float C = aB - aA - f1.x;
            const float omega = 2.0f * kPi * f1.y;
            const float dd = 2.0f * m * f1.z * omega;
            const float k = m * omega * omega;
            gamma = h * (dd + h * k);
            gamma = gamma != 0.0f ? Div(1.0f, gamma) : 0.0f;
            bias = C * h * k * gamma;
            invM += gamma;
            M.ez.z = invM != 0.0f ? Div(1.0f, invM) : 0.0f;
        } else if (K.ez.z == 0.0f) {
            M = GetInverse22(K);
        } else {
            M = GetSymInverse33(K);
        }
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
            const Vec2 P = V2(imp.x, imp.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(rA, P) + imp.z);
            vB = vB + mB * P;
            wB += iB * (Cross(rB, P) + imp.z);
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(rA.x, rA.y, rB.x, rB.y);
        c.jTmp[t + 1] = make_float4(M.ex.x, M.ex.y, M.ex.z, gamma);
        c.jTmp[t + 2] = make_float4(M.ey.x, M.ey.y, M.ey.z, bias);
        c.jTmp[t + 3] = make_float4(M.ez.x, M.ez.y, M.ez.z, 0.0f);
    } else {  // prismatic — b2PrismaticJoint.cpp:127-264
        const Vec2 d = (cB - cA) + rB - rA;
        const Vec2 localX = V2(f2.y, f2.z), localY = Cross(1.0f, localX);
        const Vec2 axis = Mul(qA, localX);
        const float a1 = Cross(d + rA, axis), a2 = Cross(rB, axis);
        float motorMass = mA + mB + iA * a1 * a1 + iB * a2 * a2;
        if (motorMass > 0.0f) motorMass = Div(1.0f, motorMass);
        const Vec2 perp = Mul(qA, localY);
        const float s1 = Cross(d + rA, perp), s2 = Cross(rB, perp);
        const float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
        const float k12 = iA * s1 + iB * s2;
        const float k13 = iA * s1 * a1 + iB * s2 * a2;
        float k22 = iA + iB;
        if (k22 == 0.0f) k22 = 1.0f;
        const float k23 = iA * a1 + iB * a2;
        const float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
        int limitState = ji.z;
        if (ji.x & kJointEnableLimit) {
            const float jointTranslation = Dot(axis, d);
            if (Abs(f1.z - f1.y) < 2.0f * kLinearSlop) {
                limitState = kLimitEqual;
            } else if (jointTranslation <= f1.y) {
                if (limitState != kLimitAtLower) {
                    limitState = kLimitAtLower;
                    imp.z = 0.0f;
                }
            } else if (jointTranslation >= f1.z) {
                if (limitState != kLimitAtUpper) {
                    limitState = kLimitAtUpper;
                    imp.z = 0.0f;
                }
            } else {
                limitState = kLimitInactive;
                imp.z = 0.0f;
            }
        } else {
            limitState = kLimitInactive;
            imp.z = 0.0f;
        }
        if (limitState != ji.z) {
            ji.z = limitState;
            c.jI[j] = ji;
        }
        if ((ji.x & kJointEnableMotor) == 0) imp.w = 0.0f;
        if (warmStarting) {
            imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
            imp.w *= dtRatio;
            const Vec2 P = imp.x * perp + (imp.w + imp.z) * axis;
            const float LA = imp.x * s1 + imp.y + (imp.w + imp.z) * a1;
            const float LB = imp.x * s2 + imp.y + (imp.w + imp.z) * a2;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        } else {
            imp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        c.jTmp[t + 0] = make_float4(axis.x, axis.y, perp.x, perp.y);
        c.jTmp[t + 1] = make_float4(a1, a2, s1, s2);
        c.jTmp[t + 2] = make_float4(k11, k12, k13, motorMass);
        c.jTmp[t + 3] = make_float4(k12, k22, k23, 0.0f);
        c.jTmp[t + 4] = make_float4(k13, k23, k33, 0.0f);
    }
    c.jImp[j] = imp;
    c.setVel(bA, make_float4(vA.x, vA.y, wA, 0.0f));
    c.setVel(bB, make_float4(vB.x, vB.y, wB, 0.0f));
}

// ---- SolveVelocityConstraints ----------------------------------------------------------------------------------
__device__ __noinline__ void JointSolveVelocity(const JointCols* cp, int j, float h) {
    const JointCols& c = *cp;
    const int4 ji = c.jI[j];
    const int type = ji.x & 0xff;
    const int bA = ji.y & 0xffff, bB = (ji.y >> 16) & 0xffff;
    const float4 msA = c.bodyMass[bA], msB = c.bodyMass[bB];
    const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
    float4 velA = c.vel(bA), velB = c.vel(bB);
    Vec2 vA = V2(velA.x, velA.y), vB = V2(velB.x, velB.y);
    float wA = velA.z, wB = velB.z;
    float4 imp = c.jImp[j];
    const size_t t = (size_t)j * kJointTmp;
    const int limitState = ji.z;

    if (type == B2G_JOINT_GEAR) {  // b2GearJoint.cpp:250-285
        const int bC = ji.w & 0xffff, bD = (ji.w >> 16) & 0xffff;
        const float4 msC = c.bodyMass[bC], msD = c.bodyMass[bD];
        const float mC = msC.x, iC = msC.y, mD = msD.x, iD = msD.y;
        float4 velC = c.vel(bC), velD = c.vel(bD);
        Vec2 vC = V2(velC.x, velC.y), vD = V2(velD.x, velD.y);
        float wC = velC.z, wD = velD.z;
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 JvAC = V2(t0.x, t0.y), JvBD = V2(t0.z, t0.w);
        const float JwA = t1.x, JwB = t1.y, JwC = t1.z, JwD = t1.w;
        float Cdot = Dot(JvAC, vA - vC) + Dot(JvBD, vB - vD);
        Cdot += (JwA * wA - JwC * wC) + (JwB * wB - JwD * wD);
        const float impulse = -t2.x * Cdot;
        imp.x += impulse;
        vA = vA + (mA * impulse) * JvAC;
        wA += iA * impulse * JwA;
        vB = vB + (mB * impulse) * JvBD;
        wB += iB * impulse * JwB;
        vC = vC - (mC * impulse) * JvAC;
        wC -= iC * impulse * JwC;
        vD = vD - (mD * impulse) * JvBD;
        wD -= iD * impulse * JwD;
        c.jImp[j] = imp;
        c.setVel(bA, make_float4(vA.x, vA.y, wA, 0.0f));
        c.setVel(bB, make_float4(vB.x, vB.y, wB, 0.0f));
        c.setVel(bC, make_float4(vC.x, vC.y, wC, 0.0f));
        c.setVel(bD, make_float4(vD.x, vD.y, wD, 0.0f));
        return;
    }
    if (type == B2G_JOINT_MOUSE) {  // b2MouseJoint.cpp:162-186
        const float4 f1 = c.jF1[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 rB = V2(t0.x, t0.y), C = V2(t0.z, t0.w);
        const Vec2 Cdot = vB + Cross(wB, rB);
        const Vec2 rhs = -(Cdot + C + t2.x * V2(imp.x, imp.y));
        Vec2 impulse = V2(t1.x * rhs.x + t1.y * rhs.y, t1.z * rhs.x + t1.w * rhs.y);  // b2Mul(m_mass, rhs)
        const Vec2 oldImpulse = V2(imp.x, imp.y);
        imp.x += impulse.x;
        imp.y += impulse.y;
        const float maxImpulse = h * f1.x;
        if (LengthSquared(V2(imp.x, imp.y)) > maxImpulse * maxImpulse) {
            const float scale = Div(maxImpulse, Length(V2(imp.x, imp.y)));
            imp.x *= scale;
            imp.y *= scale;
        }
        impulse = V2(imp.x, imp.y) - oldImpulse;
        vB = vB + mB * impulse;
        wB += iB * Cross(rB, impulse);
    } else if (type == B2G_JOINT_REVOLUTE) {  // b2RevoluteJoint.cpp:178-277
        const float4 f1 = c.jF1[j], f2 = c.jF2[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2], t3 = c.jTmp[t + 3];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w);
        Mat33 M;
        M.ex = V3(t1.x, t1.y, t1.z); M.ey = V3(t2.x, t2.y, t2.z); M.ez = V3(t3.x, t3.y, t3.z);
        const float motorMass = t1.w;
        const bool fixedRotation = (iA + iB == 0.0f);
        if ((ji.x & kJointEnableMotor) && limitState != kLimitEqual && !fixedRotation) {
            const float Cdot = wB - wA - f2.x;
            float impulse = -motorMass * Cdot;
            const float oldImpulse = imp.w;
            const float maxImpulse = h * f1.w;
            imp.w = Clamp(imp.w + impulse, -maxImpulse, maxImpulse);
            impulse = imp.w - oldImpulse;
            wA -= iA * impulse;
            wB += iB * impulse;
        }
        if ((ji.x & kJointEnableLimit) && limitState != kLimitInactive && !fixedRotation) {
            const Vec2 Cdot1 = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            const float Cdot2 = wB - wA;
            const Vec3 s = Solve33(M, V3(Cdot1.x, Cdot1.y, Cdot2));
            Vec3 impulse = V3(-s.x, -s.y, -s.z);
            if (limitState == kLimitEqual) {
                imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
            } else if (limitState == kLimitAtLower) {
                const float newImpulse = imp.z + impulse.z;
                if (newImpulse < 0.0f) {
                    const Vec2 rhs = -Cdot1 + imp.z * V2(M.ez.x, M.ez.y);
                    const Vec2 reduced = Solve22(M.ex.x, M.ey.x, M.ex.y, M.ey.y, rhs);
                    impulse.x = reduced.x;
                    impulse.y = reduced.y;
                    impulse.z = -imp.z;
                    imp.x += reduced.x;
                    imp.y += reduced.y;
                    imp.z = 0.0f;
                } else {
                    imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
                }
            } else if (limitState == kLimitAtUpper) {
                const float newImpulse = imp.z + impulse.z;
                if (newImpulse > 0.0f) {
                    const Vec2 rhs = -Cdot1 + imp.z * V2(M.ez.x, M.ez.y);
                    const Vec2 reduced = Solve22(M.ex.x, M.ey.x, M.ex.y, M.ey.y, rhs);
                    impulse.x = reduced.x;
                    impulse.y = reduced.y;
                    impulse.z = -imp.z;
                    imp.x += reduced.x;
                    imp.y += reduced.y;
                    imp.z = 0.0f;
                } else {
                    imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
                }
            }
            const Vec2 P = V2(impulse.x, impulse.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(rA, P) + impulse.z);
            vB = vB + mB * P;
            wB += iB * (Cross(rB, P) + impulse.z);
        } else {
            const Vec2 Cdot = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            const Vec2 impulse = Solve22(M.ex.x, M.ey.x, M.ex.y, M.ey.y, -Cdot);
            imp.x += impulse.x;
            imp.y += impulse.y;
            vA = vA - mA * impulse;
            wA -= iA * Cross(rA, impulse);
            vB = vB + mB * impulse;
            wB += iB * Cross(rB, impulse);
        }
    } else if (type == B2G_JOINT_DISTANCE) {  // b2DistanceJoint.cpp:154-181
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w), u = V2(t1.x, t1.y);
        const Vec2 vpA = vA + Cross(wA, rA);
        const Vec2 vpB = vB + Cross(wB, rB);
        const float Cdot = Dot(u, vpB - vpA);
        const float impulse = -t1.z * (Cdot + t2.x + t1.w * imp.x);
        imp.x += impulse;
        const Vec2 P = impulse * u;
        vA = vA - mA * P;
        wA -= iA * Cross(rA, P);
        vB = vB + mB * P;
        wB += iB * Cross(rB, P);
    } else if (type == B2G_JOINT_PULLEY) {  // b2PulleyJoint.cpp:166-193
        const float4 f2 = c.jF2[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w), uA = V2(t1.x, t1.y), uB = V2(t1.z, t1.w);
        const Vec2 vpA = vA + Cross(wA, rA);
        const Vec2 vpB = vB + Cross(wB, rB);
        const float Cdot = -Dot(uA, vpA) - f2.x * Dot(uB, vpB);
        const float impulse = -t2.x * Cdot;
        imp.x += impulse;
        const Vec2 PA = -impulse * uA;
        const Vec2 PB = (-f2.x * impulse) * uB;
        vA = vA + mA * PA;
        wA += iA * Cross(rA, PA);
        vB = vB + mB * PB;
        wB += iB * Cross(rB, PB);
    } else if (type == B2G_JOINT_MOTOR) {  // b2MotorJoint.cpp:135-193
        const float4 f1 = c.jF1[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w);
        const float inv_h = Div(1.0f, h);
        {  // angular constraint
            const float Cdot = wB - wA + inv_h * f1.w * t2.w;
            float impulse = -t2.x * Cdot;
            const float oldImpulse = imp.z;
            const float maxImpulse = h * f1.y;
            imp.z = Clamp(imp.z + impulse, -maxImpulse, maxImpulse);
            impulse = imp.z - oldImpulse;
            wA -= iA * impulse;
            wB += iB * impulse;
        }
        {  // linear constraint
            const Vec2 Cdot = vB + Cross(wB, rB) - vA - Cross(wA, rA) + (inv_h * f1.w) * V2(t2.y, t2.z);
            Vec2 impulse = -V2(t1.x * Cdot.x + t1.y * Cdot.y, t1.z * Cdot.x + t1.w * Cdot.y);
            const Vec2 oldImpulse = V2(imp.x, imp.y);
            Vec2 li = oldImpulse + impulse;
            const float maxImpulse = h * f1.x;
            if (Dot(li, li) > maxImpulse * maxImpulse) {
                Normalize(li);
                li = maxImpulse * li;
            }
            imp.x = li.x;
            imp.y = li.y;
            impulse = li - oldImpulse;
            vA = vA - mA * impulse;
            wA -= iA * Cross(rA, impulse);
            vB = vB + mB * impulse;
            wB += iB * Cross(rB, impulse);
        }
    } else if (type == B2G_JOINT_ROPE) {  // b2RopeJoint.cpp:135-169
        const float4 f1 = c.jF1[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w), u = V2(t1.x, t1.y);
        const Vec2 vpA = vA + Cross(wA, rA);
        const Vec2 vpB = vB + Cross(wB, rB);
        const float C = t1.w - f1.x;
        float Cdot = Dot(u, vpB - vpA);
        if (C < 0.0f) Cdot += Div(1.0f, h) * C;  // predictive constraint: data.step.inv_dt * C
        float impulse = -t1.z * Cdot;
        const float oldImpulse = imp.x;
        imp.x = Min(0.0f, imp.x + impulse);
        impulse = imp.x - oldImpulse;
        const Vec2 P = impulse * u;
        vA = vA - mA * P;
        wA -= iA * Cross(rA, P);
        vB = vB + mB * P;
        wB += iB * Cross(rB, P);
    } else if (type == B2G_JOINT_FRICTION) {  // b2FrictionJoint.cpp:130-186
        const float4 f1 = c.jF1[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w);
        {  // angular friction
            const float Cdot = wB - wA;
            float impulse = -t2.x * Cdot;
            const float oldImpulse = imp.z;
            const float maxImpulse = h * f1.y;
            imp.z = Clamp(imp.z + impulse, -maxImpulse, maxImpulse);
            impulse = imp.z - oldImpulse;
            wA -= iA * impulse;
            wB += iB * impulse;
        }
        {  // linear friction
            const Vec2 Cdot = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            Vec2 impulse = -V2(t1.x * Cdot.x + t1.y * Cdot.y, t1.z * Cdot.x + t1.w * Cdot.y);  // -b2Mul(m_linearMass, Cdot)
            const Vec2 oldImpulse = V2(imp.x, imp.y);
            Vec2 li = oldImpulse + impulse;
            const float maxImpulse = h * f1.x;
            if (Dot(li, li) > maxImpulse * maxImpulse) {
                Normalize(li);
                li = maxImpulse * li;  // b2Vec2::operator*=(float): x *= a
            }
            imp.x = li.x;
            imp.y = li.y;
            impulse = li - oldImpulse;
            vA = vA - mA * impulse;
            wA -= iA * Cross(rA, impulse);
            vB = vB + mB * impulse;
            wB += iB * Cross(rB, impulse);
        }
    } else if (type == B2G_JOINT_WHEEL) {  // b2WheelJoint.cpp:201-261
        const float4 f1 = c.jF1[j], f2 = c.jF2[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2], t3 = c.jTmp[t + 3];
        const Vec2 ax = V2(t0.x, t0.y), ay = V2(t0.z, t0.w);
        const float sAx = t1.x, sBx = t1.y, sAy = t1.z, sBy = t1.w;
        {  // spring constraint
            const float Cdot = Dot(ax, vB - vA) + sBx * wB - sAx * wA;
            const float impulse = -t2.y * (Cdot + t3.x + t2.w * imp.y);
            imp.y += impulse;
            const Vec2 P = impulse * ax;
            const float LA = impulse * sAx;
            const float LB = impulse * sBx;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        }
        {  // rotational motor constraint
            const float Cdot = wB - wA - f2.x;
            float impulse = -t2.z * Cdot;
            const float oldImpulse = imp.w;
            const float maxImpulse = h * f1.w;
            imp.w = Clamp(imp.w + impulse, -maxImpulse, maxImpulse);
            impulse = imp.w - oldImpulse;
            wA -= iA * impulse;
            wB += iB * impulse;
        }
        {  // point to line constraint
            const float Cdot = Dot(ay, vB - vA) + sBy * wB - sAy * wA;
            const float impulse = -t2.x * Cdot;
            imp.x += impulse;
            const Vec2 P = impulse * ay;
            const float LA = impulse * sAy;
            const float LB = impulse * sBy;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        }
    } else if (type == B2G_JOINT_WELD) {  // b2WeldJoint.cpp:156-213
        const float4 f1 = c.jF1[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2], t3 = c.jTmp[t + 3];
        const Vec2 rA = V2(t0.x, t0.y), rB = V2(t0.z, t0.w);
        Mat33 M;
        M.ex = V3(t1.x, t1.y, t1.z); M.ey = V3(t2.x, t2.y, t2.z); M.ez = V3(t3.x, t3.y, t3.z);
        if (f1.y > 0.0f) {
            const float Cdot2 = wB - wA;
            const float impulse2 = -M.ez.z * (Cdot2 + t2.w + t1.w * imp.z);
            imp.z += impulse2;
            wA -= iA * impulse2;
            wB += iB * impulse2;
            const Vec2 Cdot1 = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            const Vec2 impulse1 = -Mul22(M, Cdot1);
            imp.x += impulse1.x;
            imp.y += impulse1.y;
            const Vec2 P = impulse1;
            vA = vA - mA * P;
            wA -= iA * Cross(rA, P);
            vB = vB + mB * P;
            wB += iB * Cross(rB, P);
        } else {
            const Vec2 Cdot1 = vB + Cross(wB, rB) - vA - Cross(wA, rA);
            const float Cdot2 = wB - wA;
            const Vec3 mm = Mul33(M, V3(Cdot1.x, Cdot1.y, Cdot2));
            const Vec3 impulse = V3(-mm.x, -mm.y, -mm.z);
            imp.x += impulse.x; imp.y += impulse.y; imp.z += impulse.z;
            const Vec2 P = V2(impulse.x, impulse.y);
            vA = vA - mA * P;
            wA -= iA * (Cross(rA, P) + impulse.z);
            vB = vB + mB * P;
            wB += iB * (Cross(rB, P) + impulse.z);
        }
    } else {  // prismatic — b2PrismaticJoint.cpp:266-372
        const float4 f1 = c.jF1[j], f2 = c.jF2[j];
        const float4 t0 = c.jTmp[t + 0], t1 = c.jTmp[t + 1], t2 = c.jTmp[t + 2], t3 = c.jTmp[t + 3], t4 = c.jTmp[t + 4];
        const Vec2 axis = V2(t0.x, t0.y), perp = V2(t0.z, t0.w);
        const float a1 = t1.x, a2 = t1.y, s1 = t1.z, s2 = t1.w;
        Mat33 K;
        K.ex = V3(t2.x, t2.y, t2.z); K.ey = V3(t3.x, t3.y, t3.z); K.ez = V3(t4.x, t4.y, t4.z);
        const float motorMass = t2.w;
        if ((ji.x & kJointEnableMotor) && limitState != kLimitEqual) {
            const float Cdot = Dot(axis, vB - vA) + a2 * wB - a1 * wA;
            float impulse = motorMass * (f2.x - Cdot);
            const float oldImpulse = imp.w;
            const float maxImpulse = h * f1.w;
            imp.w = Clamp(imp.w + impulse, -maxImpulse, maxImpulse);
            impulse = imp.w - oldImpulse;
            const Vec2 P = impulse * axis;
            const float LA = impulse * a1;
            const float LB = impulse * a2;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        }
        Vec2 Cdot1;
        Cdot1.x = Dot(perp, vB - vA) + s2 * wB - s1 * wA;
        Cdot1.y = wB - wA;
        if ((ji.x & kJointEnableLimit) && limitState != kLimitInactive) {
            const float Cdot2 = Dot(axis, vB - vA) + a2 * wB - a1 * wA;
            const Vec3 f1v = V3(imp.x, imp.y, imp.z);
            Vec3 df = Solve33(K, V3(-Cdot1.x, -Cdot1.y, -Cdot2));
            imp.x += df.x; imp.y += df.y; imp.z += df.z;
            if (limitState == kLimitAtLower) imp.z = Max(imp.z, 0.0f);
            else if (limitState == kLimitAtUpper) imp.z = Min(imp.z, 0.0f);
            const Vec2 b = -Cdot1 - (imp.z - f1v.z) * V2(K.ez.x, K.ez.y);
            const Vec2 f2r = Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, b) + V2(f1v.x, f1v.y);
            imp.x = f2r.x;
            imp.y = f2r.y;
            df = V3(imp.x - f1v.x, imp.y - f1v.y, imp.z - f1v.z);
            const Vec2 P = df.x * perp + df.z * axis;
            const float LA = df.x * s1 + df.y + df.z * a1;
            const float LB = df.x * s2 + df.y + df.z * a2;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        } else {
            const Vec2 df = Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, -Cdot1);
            imp.x += df.x;
            imp.y += df.y;
            const Vec2 P = df.x * perp;
            const float LA = df.x * s1 + df.y;
            const float LB = df.x * s2 + df.y;
            vA = vA - mA * P;
            wA -= iA * LA;
            vB = vB + mB * P;
            wB += iB * LB;
        }
    }
    c.jImp[j] = imp;
    c.setVel(bA, make_float4(vA.x, vA.y, wA, 0.0f));
    c.setVel(bB, make_float4(vB.x, vB.y, wB, 0.0f));
}

// ---- SolvePositionConstraints; returns the joint's "okay" -------------------------------------------------------------
__device__ __noinline__ bool JointSolvePosition(const JointCols* cp, int j) {
    const JointCols& c = *cp;
    const int4 ji = c.jI[j];
    const int type = ji.x & 0xff;
    const int bA = ji.y & 0xffff, bB = (ji.y >> 16) & 0xffff;
    const float4 f0 = c.jF0[j], f1 = c.jF1[j];
    const float4 msA = c.bodyMass[bA], msB = c.bodyMass[bB];
    const float mA = msA.x, iA = msA.y, mB = msB.x, iB = msB.y;
    float4 swA = c.sweep(bA), swB = c.sweep(bB);
    Vec2 cA = V2(swA.x, swA.y), cB = V2(swB.x, swB.y);
    float aA = swA.z, aB = swB.z;
    const Vec2 lA = V2(f0.x, f0.y) - V2(msA.z, msA.w), lB = V2(f0.z, f0.w) - V2(msB.z, msB.w);
    const int limitState = ji.z;
    bool okay;

    if (type == B2G_JOINT_MOUSE) return true;  // b2MouseJoint.cpp:188-192
    if (type == B2G_JOINT_GEAR) {  // b2GearJoint.cpp:287-383
        const int bC = ji.w & 0xffff, bD = (ji.w >> 16) & 0xffff;
        const int typeA = (ji.x >> 11) & 3, typeB = (ji.x >> 13) & 3;
        const float4 f2 = c.jF2[j], imp = c.jImp[j];
        const float4 msC = c.bodyMass[bC], msD = c.bodyMass[bD];
        const float mC = msC.x, iC = msC.y, mD = msD.x, iD = msD.y;
        float4 swC = c.sweep(bC), swD = c.sweep(bD);
        Vec2 cC = V2(swC.x, swC.y), cD = V2(swD.x, swD.y);
        float aC = swC.z, aD = swD.z;
        const Rot qA = MakeRot(aA), qB = MakeRot(aB), qC = MakeRot(aC), qD = MakeRot(aD);
        const float ratio = f1.z, constant = f1.w;
        const float linearError = 0.0f;
        float coordinateA, coordinateB;
        Vec2 JvAC, JvBD;
        float JwA, JwB, JwC, JwD;
        float mass = 0.0f;
        if (typeA == B2G_JOINT_REVOLUTE) {
            JvAC = V2(0.0f, 0.0f);
            JwA = 1.0f;
            JwC = 1.0f;
            mass += iA + iC;
            coordinateA = aA - aC - f0.x;
        } else {
            const Vec2 axisC = V2(f1.x, f1.y);
            const Vec2 u = Mul(qC, axisC);
            const Vec2 rC = Mul(qC, V2(f0.x, f0.y) - V2(msC.z, msC.w));
            const Vec2 rA = Mul(qA, V2(f0.z, f0.w) - V2(msA.z, msA.w));
            JvAC = u;
            JwC = Cross(rC, u);
            JwA = Cross(rA, u);
            mass += mC + mA + iC * JwC * JwC + iA * JwA * JwA;
            const Vec2 pC = V2(f0.x, f0.y) - V2(msC.z, msC.w);
            const Vec2 pA = MulT(qC, rA + (cA - cC));
            coordinateA = Dot(pA - pC, axisC);
        }
        if (typeB == B2G_JOINT_REVOLUTE) {
            JvBD = V2(0.0f, 0.0f);
            JwB = ratio;
            JwD = ratio;
            mass += ratio * ratio * (iB + iD);
            coordinateB = aB - aD - f2.x;
        } else {
            const Vec2 axisD = V2(imp.z, imp.w);
            const Vec2 u = Mul(qD, axisD);
            const Vec2 rD = Mul(qD, V2(f2.x, f2.y) - V2(msD.z, msD.w));
            const Vec2 rB = Mul(qB, V2(f2.z, f2.w) - V2(msB.z, msB.w));
            JvBD = ratio * u;
            JwD = ratio * Cross(rD, u);
            JwB = ratio * Cross(rB, u);
            mass += ratio * ratio * (mD + mB) + iD * JwD * JwD + iB * JwB * JwB;
            const Vec2 pD = V2(f2.x, f2.y) - V2(msD.z, msD.w);
            const Vec2 pB = MulT(qD, rB + (cB - cD));
            coordinateB = Dot(pB - pD, axisD);
        }
        const float C = (coordinateA + ratio * coordinateB) - constant;
        float impulse = 0.0f;
        if (mass > 0.0f) impulse = Div(-C, mass);
        cA = cA + (mA * impulse) * JvAC;
        aA += iA * impulse * JwA;
        cB = cB + (mB * impulse) * JvBD;
        aB += iB * impulse * JwB;
        cC = cC - (mC * impulse) * JvAC;
        aC -= iC * impulse * JwC;
        cD = cD - (mD * impulse) * JvBD;
        aD -= iD * impulse * JwD;
        c.setSweep(bA, make_float4(cA.x, cA.y, aA, swA.w));
        c.setSweep(bB, make_float4(cB.x, cB.y, aB, swB.w));
        c.setSweep(bC, make_float4(cC.x, cC.y, aC, swC.w));
        c.setSweep(bD, make_float4(cD.x, cD.y, aD, swD.w));
        return linearError < kLinearSlop;
    }
    if (type == B2G_JOINT_REVOLUTE) {  // b2RevoluteJoint.cpp:279-371
        float angularError = 0.0f, positionError = 0.0f;
        const bool fixedRotation = (iA + iB == 0.0f);
        if ((ji.x & kJointEnableLimit) && limitState != kLimitInactive && !fixedRotation) {
            const float motorMass = c.jTmp[(size_t)j * kJointTmp + 1].w;
            const float angle = aB - aA - f1.x;
            float limitImpulse = 0.0f;
            if (limitState == kLimitEqual) {
                const float C = Clamp(angle - f1.y, -kMaxAngularCorrection, kMaxAngularCorrection);
                limitImpulse = -motorMass * C;
                angularError = Abs(C);
            } else if (limitState == kLimitAtLower) {
                float C = angle - f1.y;
                angularError = -C;
                C = Clamp(C + kAngularSlop, -kMaxAngularCorrection, 0.0f);
                limitImpulse = -motorMass * C;
            } else if (limitState == kLimitAtUpper) {
                float C = angle - f1.z;
                angularError = C;
                C = Clamp(C - kAngularSlop, 0.0f, kMaxAngularCorrection);
                limitImpulse = -motorMass * C;
            }
            aA -= iA * limitImpulse;
            aB += iB * limitImpulse;
        }
        {
            const Rot qA = MakeRot(aA), qB = MakeRot(aB);
            const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
            const Vec2 C = cB + rB - cA - rA;
            positionError = Length(C);
            const float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
            const float k12 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
            const float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
            const Vec2 s = Solve22(k11, k12, k12, k22, C);
            const Vec2 impulse = -s;
            cA = cA - mA * impulse;
            aA -= iA * Cross(rA, impulse);
            cB = cB + mB * impulse;
            aB += iB * Cross(rB, impulse);
        }
        okay = positionError <= kLinearSlop && angularError <= kAngularSlop;
    } else if (type == B2G_JOINT_DISTANCE) {  // b2DistanceJoint.cpp:183-220
        if (f1.y > 0.0f) return true;  // soft constraint: no position correction
        const float mass = c.jTmp[(size_t)j * kJointTmp + 1].z;
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        Vec2 u = cB + rB - cA - rA;
        const float length = Normalize(u);
        float C = length - f1.x;
        C = Clamp(C, -kMaxLinearCorrection, kMaxLinearCorrection);
        const float impulse = -mass * C;
        const Vec2 P = impulse * u;
        cA = cA - mA * P;
        aA -= iA * Cross(rA, P);
        cB = cB + mB * P;
        aB += iB * Cross(rB, P);
        okay = Abs(C) < kLinearSlop;
    } else if (type == B2G_JOINT_PULLEY) {  // b2PulleyJoint.cpp:195-264
        const float4 f2 = c.jF2[j];
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        Vec2 uA = cA + rA - V2(f1.x, f1.y);
        Vec2 uB = cB + rB - V2(f1.z, f1.w);
        const float lengthA = Length(uA), lengthB = Length(uB);
        if (lengthA > 10.0f * kLinearSlop) uA = Div(1.0f, lengthA) * uA; else uA = V2(0.0f, 0.0f);
        if (lengthB > 10.0f * kLinearSlop) uB = Div(1.0f, lengthB) * uB; else uB = V2(0.0f, 0.0f);
        const float ruA = Cross(rA, uA), ruB = Cross(rB, uB);
        const float pmA = mA + iA * ruA * ruA;
        const float pmB = mB + iB * ruB * ruB;
        float mass = pmA + f2.x * f2.x * pmB;
        if (mass > 0.0f) mass = Div(1.0f, mass);
        const float C = f2.y - lengthA - f2.x * lengthB;
        const float linearError = Abs(C);
        const float impulse = -mass * C;
        const Vec2 PA = -impulse * uA;
        const Vec2 PB = (-f2.x * impulse) * uB;
        cA = cA + mA * PA;
        aA += iA * Cross(rA, PA);
        cB = cB + mB * PB;
        aB += iB * Cross(rB, PB);
        okay = linearError < kLinearSlop;
    } else if (type == B2G_JOINT_ROPE) {  // b2RopeJoint.cpp:171-206
        const float mass = c.jTmp[(size_t)j * kJointTmp + 1].z;
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        Vec2 u = cB + rB - cA - rA;
        const float length = Normalize(u);
        float C = length - f1.x;
        C = Clamp(C, 0.0f, kMaxLinearCorrection);
        const float impulse = -mass * C;
        const Vec2 P = impulse * u;
        cA = cA - mA * P;
        aA -= iA * Cross(rA, P);
        cB = cB + mB * P;
        aB += iB * Cross(rB, P);
        okay = length - f1.x < kLinearSlop;
    } else if (type == B2G_JOINT_FRICTION || type == B2G_JOINT_MOTOR) {  // b2FrictionJoint.cpp:188-193, b2MotorJoint.cpp:195-200
        return true;  // no position correction
    } else if (type == B2G_JOINT_WHEEL) {  // b2WheelJoint.cpp:263-312
        const float4 f2 = c.jF2[j];
        const float4 t1 = c.jTmp[(size_t)j * kJointTmp + 1];
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        const Vec2 d = (cB - cA) + rB - rA;
        const Vec2 ay = Mul(qA, Cross(1.0f, V2(f2.y, f2.z)));
        const float sAy = Cross(d + rA, ay), sBy = Cross(rB, ay);
        const float C = Dot(d, ay);
        // the reference evaluates k with the members m_sAy / m_sBy of InitVelocityConstraints, not the locals
        const float k = mA + mB + iA * t1.z * t1.z + iB * t1.w * t1.w;
        const float impulse = k != 0.0f ? Div(-C, k) : 0.0f;
        const Vec2 P = impulse * ay;
        const float LA = impulse * sAy;
        const float LB = impulse * sBy;
        cA = cA - mA * P;
        aA -= iA * LA;
        cB = cB + mB * P;
        aB += iB * LB;
        okay = Abs(C) <= kLinearSlop;
    } else if (type == B2G_JOINT_WELD) {  // b2WeldJoint.cpp:215-289
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        float positionError, angularError;
        Mat33 K;
        K.ex.x = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        K.ey.x = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        K.ez.x = -rA.y * iA - rB.y * iB;
        K.ex.y = K.ey.x;
        K.ey.y = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        K.ez.y = rA.x * iA + rB.x * iB;
        K.ex.z = K.ez.x;
        K.ey.z = K.ez.y;
        K.ez.z = iA + iB;
        if (f1.y > 0.0f) {
            const Vec2 C1 = cB + rB - cA - rA;
            positionError = Length(C1);
            angularError = 0.0f;
            const Vec2 P = -Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, C1);
            cA = cA - mA * P;
            aA -= iA * Cross(rA, P);
            cB = cB + mB * P;
            aB += iB * Cross(rB, P);
        } else {
            const Vec2 C1 = cB + rB - cA - rA;
            const float C2 = aB - aA - f1.x;
            positionError = Length(C1);
            angularError = Abs(C2);
            Vec3 impulse;
            if (K.ez.z > 0.0f) {
                const Vec3 s3 = Solve33(K, V3(C1.x, C1.y, C2));
                impulse = V3(-s3.x, -s3.y, -s3.z);
            } else {
                const Vec2 impulse2 = -Solve22(K.ex.x, K.ey.x, K.ex.y, K.ey.y, C1);
                impulse = V3(impulse2.x, impulse2.y, 0.0f);
            }
            const Vec2 P = V2(impulse.x, impulse.y);
            cA = cA - mA * P;
            aA -= iA * (Cross(rA, P) + impulse.z);
            cB = cB + mB * P;
            aB += iB * (Cross(rB, P) + impulse.z);
        }
        okay = positionError <= kLinearSlop && angularError <= kAngularSlop;
    } else {  // prismatic — b2PrismaticJoint.cpp:374-491
        const float4 f2 = c.jF2[j];
        const Rot qA = MakeRot(aA), qB = MakeRot(aB);
        const Vec2 rA = Mul(qA, lA), rB = Mul(qB, lB);
        const Vec2 d = cB + rB - cA - rA;
        const Vec2 localX = V2(f2.y, f2.z), localY = Cross(1.0f, localX);
        const Vec2 axis = Mul(qA, localX);
        const float a1 = Cross(d + rA, axis), a2 = Cross(rB, axis);
        const Vec2 perp = Mul(qA, localY);
        const float s1 = Cross(d + rA, perp), s2 = Cross(rB, perp);
        Vec3 impulse;
        Vec2 C1;
        C1.x = Dot(perp, d);
        C1.y = aB - aA - f1.x;
        float linearError = Abs(C1.x);
        const float angularError = Abs(C1.y);
        bool active = false;
        float C2 = 0.0f;
        if (ji.x & kJointEnableLimit) {
            const float translation = Dot(axis, d);
            if (Abs(f1.z - f1.y) < 2.0f * kLinearSlop) {
                C2 = Clamp(translation, -kMaxLinearCorrection, kMaxLinearCorrection);
                linearError = Max(linearError, Abs(translation));
                active = true;
            } else if (translation <= f1.y) {
                C2 = Clamp(translation - f1.y + kLinearSlop, -kMaxLinearCorrection, 0.0f);
                linearError = Max(linearError, f1.y - translation);
                active = true;
            } else if (translation >= f1.z) {
                C2 = Clamp(translation - f1.z - kLinearSlop, 0.0f, kMaxLinearCorrection);
                linearError = Max(linearError, translation - f1.z);
                active = true;
            }
        }
        if (active) {
            const float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
            const float k12 = iA * s1 + iB * s2;
            const float k13 = iA * s1 * a1 + iB * s2 * a2;
            float k22 = iA + iB;
            if (k22 == 0.0f) k22 = 1.0f;
            const float k23 = iA * a1 + iB * a2;
            const float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
            Mat33 K;
            K.ex = V3(k11, k12, k13); K.ey = V3(k12, k22, k23); K.ez = V3(k13, k23, k33);
            impulse = Solve33(K, V3(-C1.x, -C1.y, -C2));
        } else {
            const float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
            const float k12 = iA * s1 + iB * s2;
            float k22 = iA + iB;
            if (k22 == 0.0f) k22 = 1.0f;
            const Vec2 impulse1 = Solve22(k11, k12, k12, k22, -C1);
            impulse = V3(impulse1.x, impulse1.y, 0.0f);
        }
        const Vec2 P = impulse.x * perp + impulse.z * axis;
        const float LA = impulse.x * s1 + impulse.y + impulse.z * a1;
        const float LB = impulse.x * s2 + impulse.y + impulse.z * a2;
        cA = cA - mA * P;
        aA -= iA * LA;
        cB = cB + mB * P;
        aB += iB * LB;
        okay = linearError <= kLinearSlop && angularError <= kAngularSlop;
    }
    c.setSweep(bA, make_float4(cA.x, cA.y, aA, swA.w));
    c.setSweep(bB, make_float4(cB.x, cB.y, aB, swB.w));
    return okay;
}
