// fp32 vector / quaternion helpers for the sm_100a kernels.
//
// Numerical contract: the translation units that include this header are compiled with
// -fmad=false, so the ONLY fused multiply-adds are the explicit fmaf() calls below.  That
// makes every result a pure function of IEEE-754 single-precision operations (+,-,*,/,sqrt,
// fma, rint) whose order is fixed by the source, which is what lets the state-validity
// booleans be compared bit-for-bit with the CPU oracle.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MOPA_HD __host__ __device__ __forceinline__
#define MOPA_HD_COLD __host__ __device__ __noinline__   // rare paths kept out of line (register pressure of the callers)
#else
#define MOPA_HD inline
#define MOPA_HD_COLD inline
#endif

namespace mopa {

struct V3 { float x, y, z; };

MOPA_HD float dot(const V3 &a, const V3 &b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
MOPA_HD V3 cross(const V3 &a, const V3 &b) {
    return V3{fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}
MOPA_HD V3 operator-(const V3 &a, const V3 &b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
MOPA_HD V3 operator+(const V3 &a, const V3 &b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
MOPA_HD V3 neg(const V3 &a) { return V3{-a.x, -a.y, -a.z}; }
MOPA_HD float len(const V3 &a) { return sqrtf(dot(a, a)); }
// a + s*b, one fma per component
MOPA_HD V3 madd(const V3 &a, float s, const V3 &b) { return V3{fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z)}; }

// row-major 3x3
struct M3 { float m[9]; };
MOPA_HD V3 mulMV(const M3 &M, const V3 &v) {
    return V3{fmaf(M.m[2], v.z, fmaf(M.m[1], v.y, M.m[0] * v.x)), fmaf(M.m[5], v.z, fmaf(M.m[4], v.y, M.m[3] * v.x)),
              fmaf(M.m[8], v.z, fmaf(M.m[7], v.y, M.m[6] * v.x))};
}
MOPA_HD V3 mulMTV(const M3 &M, const V3 &v) {
    return V3{fmaf(M.m[6], v.z, fmaf(M.m[3], v.y, M.m[0] * v.x)), fmaf(M.m[7], v.z, fmaf(M.m[4], v.y, M.m[1] * v.x)),
              fmaf(M.m[8], v.z, fmaf(M.m[5], v.y, M.m[2] * v.x))};
}
MOPA_HD M3 mulMM(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C.m[3 * i + j] = fmaf(A.m[3 * i + 2], B.m[6 + j], fmaf(A.m[3 * i + 1], B.m[3 + j], A.m[3 * i] * B.m[j]));
    return C;
}
MOPA_HD V3 col(const M3 &M, int k) { return V3{M.m[k], M.m[3 + k], M.m[6 + k]}; }

struct Q4 { float w, x, y, z; };
MOPA_HD Q4 qmul(const Q4 &a, const Q4 &b) {
    Q4 q;
    q.w = fmaf(-a.z, b.z, fmaf(-a.y, b.y, fmaf(-a.x, b.x, a.w * b.w)));
    q.x = fmaf(-a.z, b.y, fmaf(a.y, b.z, fmaf(a.x, b.w, a.w * b.x)));
    q.y = fmaf(a.z, b.x, fmaf(a.y, b.w, fmaf(-a.x, b.z, a.w * b.y)));
    q.z = fmaf(a.z, b.w, fmaf(-a.y, b.x, fmaf(a.x, b.y, a.w * b.z)));
    return q;
}
MOPA_HD M3 q2m(const Q4 &q) {
    float ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    float xy = q.x * q.y, wz = q.w * q.z, xz = q.x * q.z, wy = q.w * q.y, yz = q.y * q.z, wx = q.w * q.x;
    M3 M;
    M.m[0] = ((ww + xx) - yy) - zz;
    M.m[1] = (xy - wz) + (xy - wz);
    M.m[2] = (xz + wy) + (xz + wy);
    M.m[3] = (xy + wz) + (xy + wz);
    M.m[4] = ((ww - xx) + yy) - zz;
    M.m[5] = (yz - wx) + (yz - wx);
    M.m[6] = (xz - wy) + (xz - wy);
    M.m[7] = (yz + wx) + (yz + wx);
    M.m[8] = ((ww - xx) - yy) + zz;
    return M;
}

// sin/cos by Cody-Waite reduction to [-pi/4, pi/4] and minimax polynomials (|x| < ~100).
MOPA_HD void sincos_cw(float x, float *s, float *c) {
    float k = rintf(x * 0.6366197466850281f);
    int q = (int)k;
    float r = fmaf(-k, 1.5707963705062866f, x);
    r = fmaf(-k, -4.371138828673793e-08f, r);
    float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    float sn = fmaf(r * r2, ps, r);
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    float cs = fmaf(r2 * r2, pc, fmaf(r2, -0.5f, 1.0f));
    float ss = (q & 1) ? cs : sn;
    float cc = (q & 1) ? sn : cs;
    if (q & 2) ss = -ss;
    if ((q + 1) & 2) cc = -cc;
    *s = ss;
    *c = cc;
}

}  // namespace mopa
