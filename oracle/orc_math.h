/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * Arithmetic conventions shared by all oracle translation units.  The oracle is
 * compiled twice: ORC_F32 (float, the arithmetic the sm_100a kernels use; built with
 * -ffp-contract=off so that the only fused operations are the explicit MAD() calls,
 * which the kernels mirror with __fmaf_rn) and ORC_F64 (double, libm trig; an
 * independent higher-precision statement used to count fp32-induced flips).
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <stdint.h>

#ifdef ORC_F64
typedef double R;
#define MAD(a, b, c) fma((a), (b), (c))
#define RSQRT_(x) sqrt(x)
#define FABS_(x) fabs(x)
#define FMIN_(a, b) fmin((a), (b))
#define FMAX_(a, b) fmax((a), (b))
#define RC(x) (x)
#else
typedef float R;
#define MAD(a, b, c) fmaf((a), (b), (c))
#define RSQRT_(x) sqrtf(x)
#define FABS_(x) fabsf(x)
#define FMIN_(a, b) fminf((a), (b))
#define FMAX_(a, b) fmaxf((a), (b))
#define RC(x) (x##f)
#endif

#define ORC_BIG RC(1.0e10)

static inline R dot3(const R *a, const R *b) { return MAD(a[2], b[2], MAD(a[1], b[1], a[0] * b[0])); }
static inline void cross3(R *c, const R *a, const R *b) {
    c[0] = MAD(a[1], b[2], -(a[2] * b[1]));
    c[1] = MAD(a[2], b[0], -(a[0] * b[2]));
    c[2] = MAD(a[0], b[1], -(a[1] * b[0]));
}
static inline void sub3(R *c, const R *a, const R *b) { c[0] = a[0] - b[0]; c[1] = a[1] - b[1]; c[2] = a[2] - b[2]; }
static inline void add3(R *c, const R *a, const R *b) { c[0] = a[0] + b[0]; c[1] = a[1] + b[1]; c[2] = a[2] + b[2]; }
static inline void scl3(R *c, const R *a, R s) { c[0] = a[0] * s; c[1] = a[1] * s; c[2] = a[2] * s; }
static inline void cpy3(R *c, const R *a) { c[0] = a[0]; c[1] = a[1]; c[2] = a[2]; }
static inline R len3(const R *a) { return RSQRT_(dot3(a, a)); }
/* c = M v  (row-major 3x3) */
static inline void mulMV(R *c, const R *M, const R *v) { c[0] = dot3(M, v); c[1] = dot3(M + 3, v); c[2] = dot3(M + 6, v); }
/* c = M^T v */
static inline void mulMTV(R *c, const R *M, const R *v) {
    c[0] = MAD(M[6], v[2], MAD(M[3], v[1], M[0] * v[0]));
    c[1] = MAD(M[7], v[2], MAD(M[4], v[1], M[1] * v[0]));
    c[2] = MAD(M[8], v[2], MAD(M[5], v[1], M[2] * v[0]));
}
/* C = A B */
static inline void mulMM(R *C, const R *A, const R *B) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = MAD(A[3 * i + 2], B[6 + j], MAD(A[3 * i + 1], B[3 + j], A[3 * i] * B[j]));
}
/* q = a (x) b, wxyz */
static inline void qmul(R *q, const R *a, const R *b) {
    R w = MAD(-a[3], b[3], MAD(-a[2], b[2], MAD(-a[1], b[1], a[0] * b[0])));
    R x = MAD(-a[3], b[2], MAD(a[2], b[3], MAD(a[1], b[0], a[0] * b[1])));
    R y = MAD(a[3], b[1], MAD(a[2], b[0], MAD(-a[1], b[3], a[0] * b[2])));
    R z = MAD(a[3], b[0], MAD(-a[2], b[1], MAD(a[1], b[2], a[0] * b[3])));
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
static inline void q2m(R *M, const R *q) {
    R w = q[0], x = q[1], y = q[2], z = q[3];
    R ww = w * w, xx = x * x, yy = y * y, zz = z * z;
    R xy = x * y, wz = w * z, xz = x * z, wy = w * y, yz = y * z, wx = w * x;
    M[0] = ((ww + xx) - yy) - zz;
    M[1] = (xy - wz) + (xy - wz);
    M[2] = (xz + wy) + (xz + wy);
    M[3] = (xy + wz) + (xy + wz);
    M[4] = ((ww - xx) + yy) - zz;
    M[5] = (yz - wx) + (yz - wx);
    M[6] = (xz - wy) + (xz - wy);
    M[7] = (yz + wx) + (yz + wx);
    M[8] = ((ww - xx) - yy) + zz;
}

/* sin/cos.  F32: Cody-Waite reduction to [-pi/4, pi/4] + minimax polynomials, spelled
 * out so the CUDA kernels can reproduce it bit for bit.  F64: libm. */
static inline void sincos_r(R x, R *s, R *c) {
#ifdef ORC_F64
    *s = sin(x);
    *c = cos(x);
#else
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
    float ss, cc;
    if (q & 1) { ss = cs; cc = sn; } else { ss = sn; cc = cs; }
    if (q & 2) ss = -ss;
    if ((q + 1) & 2) cc = -cc;
    *s = ss;
    *c = cc;
#endif
}
#endif
