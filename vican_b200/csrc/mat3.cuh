// 3x3 fp64 helpers shared by every kernel (registers only, fully unrolled).
//
// svd3(): one-sided (Hestenes) Jacobi SVD -- high relative accuracy, no LAPACK.
// It supplies the three per-node factors the reference obtains from
// np.linalg.svd (vican/bipgo.py:306-312, :323-329; vican/geometry.py:189-190):
//   rot  = U diag(1,1,det(U V^T)) V^T        (nearest rotation, det fix on the
//                                             smallest singular direction)
//   spos = U S U^T  = (M M^T)^{1/2}           (Lambda_C update)
//   sinv = U S^-1 U^T = (M M^T)^{-1/2}        (Lambda_T update)
// All three are unique functions of M (no sign/order ambiguity of the SVD leaks
// out), so any accurate SVD reproduces LAPACK's gesdd to ~cond*eps.
//
// Everything is __host__ __device__ so the same code is unit-tested on the CPU
// (tests/test_host_math.py builds it with the host compiler).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD inline
#endif

namespace vb {

// C = A * B
VB_HD void mm3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A^T * B
VB_HD void mtm3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
// C = A * B^T
VB_HD void mmt3(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
VB_HD void mv3(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
VB_HD void mtv3(const double* A, const double* x, double* y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
VB_HD double det3(const double* A) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
           A[2] * (A[3] * A[7] - A[4] * A[6]);
}
// inverse by adjugate; returns det
VB_HD double inv3(const double* A, double* I) {
    const double c00 = A[4] * A[8] - A[5] * A[7];
    const double c01 = A[5] * A[6] - A[3] * A[8];
    const double c02 = A[3] * A[7] - A[4] * A[6];
    const double d = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double r = 1.0 / d;
    I[0] = c00 * r;
    I[1] = (A[2] * A[7] - A[1] * A[8]) * r;
    I[2] = (A[1] * A[5] - A[2] * A[4]) * r;
    I[3] = c01 * r;
    I[4] = (A[0] * A[8] - A[2] * A[6]) * r;
    I[5] = (A[2] * A[3] - A[0] * A[5]) * r;
    I[6] = c02 * r;
    I[7] = (A[1] * A[6] - A[0] * A[7]) * r;
    I[8] = (A[0] * A[4] - A[1] * A[3]) * r;
    return d;
}

// one Hestenes rotation on columns p,q of G (3x3, row-major) and V
#define VB_JROT(p, q)                                                                         \
    {                                                                                         \
        const double a = G[p] * G[p] + G[3 + p] * G[3 + p] + G[6 + p] * G[6 + p];             \
        const double b = G[q] * G[q] + G[3 + q] * G[3 + q] + G[6 + q] * G[6 + q];             \
        const double g = G[p] * G[q] + G[3 + p] * G[3 + q] + G[6 + p] * G[6 + q];             \
        const double lim = 1e-17 * sqrt(a * b);                                               \
        if (fabs(g) > lim && fabs(g) > 1e-300) {                                              \
            rotated = true;                                                                   \
            const double z = (b - a) / (2.0 * g);                                             \
            const double t = (z >= 0.0 ? 1.0 : -1.0) / (fabs(z) + sqrt(1.0 + z * z));         \
            const double c = 1.0 / sqrt(1.0 + t * t);                                         \
            const double s = c * t;                                                           \
            _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                   \
                const double gp = G[3 * r + p], gq = G[3 * r + q];                            \
                G[3 * r + p] = c * gp - s * gq;                                               \
                G[3 * r + q] = s * gp + c * gq;                                               \
                const double vp = V[3 * r + p], vq = V[3 * r + q];                            \
                V[3 * r + p] = c * vp - s * vq;                                               \
                V[3 * r + q] = s * vp + c * vq;                                               \
            }                                                                                 \
        }                                                                                     \
    }

#define VB_SWAPCOL(M, p, q)                                  \
    {                                                        \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {      \
            const double tmp = M[3 * r + p];                 \
            M[3 * r + p] = M[3 * r + q];                     \
            M[3 * r + q] = tmp;                              \
        }                                                    \
    }

// M = U diag(S) V^T, S descending.  U, V orthogonal (det may be -1).
VB_HD void svd3(const double* M, double* U, double* S, double* V) {
    double G[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = M[i];
    V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 0; V[4] = 1; V[5] = 0; V[6] = 0; V[7] = 0; V[8] = 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        VB_JROT(0, 1)
        VB_JROT(0, 2)
        VB_JROT(1, 2)
        if (!rotated) break;
    }
    double s0 = sqrt(G[0] * G[0] + G[3] * G[3] + G[6] * G[6]);
    double s1 = sqrt(G[1] * G[1] + G[4] * G[4] + G[7] * G[7]);
    double s2 = sqrt(G[2] * G[2] + G[5] * G[5] + G[8] * G[8]);
    // sort descending (3-element network), permuting columns of G and V
    if (s0 < s1) { VB_SWAPCOL(G, 0, 1) VB_SWAPCOL(V, 0, 1) double t = s0; s0 = s1; s1 = t; }
    if (s0 < s2) { VB_SWAPCOL(G, 0, 2) VB_SWAPCOL(V, 0, 2) double t = s0; s0 = s2; s2 = t; }
    if (s1 < s2) { VB_SWAPCOL(G, 1, 2) VB_SWAPCOL(V, 1, 2) double t = s1; s1 = s2; s2 = t; }
    S[0] = s0; S[1] = s1; S[2] = s2;
    const double tiny = 1e-300;
    const double r0 = s0 > tiny ? 1.0 / s0 : 0.0;
    const double r1 = s1 > tiny ? 1.0 / s1 : 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { U[3 * r] = G[3 * r] * r0; U[3 * r + 1] = G[3 * r + 1] * r1; }
    if (s0 <= tiny) { U[0] = 1; U[3] = 0; U[6] = 0; }
    if (s1 <= tiny) {  // any unit vector orthogonal to u0
        const double ax = fabs(U[0]), ay = fabs(U[3]), az = fabs(U[6]);
        double e0 = 0, e1 = 0, e2 = 0;
        if (ax <= ay && ax <= az) e0 = 1; else if (ay <= az) e1 = 1; else e2 = 1;
        double x = U[3] * e2 - U[6] * e1, y = U[6] * e0 - U[0] * e2, z = U[0] * e1 - U[3] * e0;
        const double n = 1.0 / sqrt(x * x + y * y + z * z);
        U[1] = x * n; U[4] = y * n; U[7] = z * n;
    }
    if (s2 > 1e-14 * s0 && s2 > tiny) {
        const double r2 = 1.0 / s2;
#pragma unroll
        for (int r = 0; r < 3; ++r) U[3 * r + 2] = G[3 * r + 2] * r2;
    } else {  // rank deficient: complete the basis (sign irrelevant for every consumer)
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

// The three per-node factors (see file header).  Any output pointer may be null.
VB_HD void svd3_factors(const double* M, double* rot, double* spos, double* sinv) {
    double U[9], S[3], V[9];
    svd3(M, U, S, V);
    if (rot) {
        const double d = (det3(U) * det3(V) < 0.0) ? -1.0 : 1.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                rot[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + d * U[3 * i + 2] * V[3 * j + 2];
    }
    if (spos) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                spos[3 * i + j] = S[0] * U[3 * i] * U[3 * j] + S[1] * U[3 * i + 1] * U[3 * j + 1] +
                                  S[2] * U[3 * i + 2] * U[3 * j + 2];
    }
    if (sinv) {
        const double i0 = 1.0 / S[0], i1 = 1.0 / S[1], i2 = 1.0 / S[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                sinv[3 * i + j] = i0 * U[3 * i] * U[3 * j] + i1 * U[3 * i + 1] * U[3 * j + 1] +
                                  i2 * U[3 * i + 2] * U[3 * j + 2];
    }
}

// cofactor matrix C = det(X) X^{-T}; returns det(X)
VB_HD double cof3(const double* X, double* C) {
    C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
    return X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
}
VB_HD double frob2_3(const double* A) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) s += A[i] * A[i];
    return s;
}

// Same three factors as svd3_factors, computed WITHOUT an SVD when M is comfortably non-singular
// with positive determinant (the normal case: M is a weighted sum of nearly consistent rotations):
//   rot  = polar factor of M by scaled Newton iteration  X <- (g X + X^-T / g) / 2   (quadratic,
//          backward stable; 6-9 iterations of one 3x3 cofactor matrix each)
//   spos = sym(M rot^T) = (M M^T)^{1/2},   sinv = spos^-1 = (M M^T)^{-1/2}
// All three are unique functions of M, so they agree with the SVD route to ~cond * eps.  About 5x
// fewer fp64 instructions than the Jacobi SVD (no sqrt/div chains per rotation).  det(M) <= 0,
// near-singular M or a stalled iteration fall back to the SVD route (det fix needs singular vectors).
VB_HD void node_factors(const double* M, double* rot, double* spos, double* sinv) {
    const double nf2 = frob2_3(M);
    const double d0 = det3(M);
    bool ok = (nf2 > 1e-280) && (d0 > 0.0);
    double X[9];
    if (ok) {
        const double sc = 1.0 / sqrt(nf2);
#pragma unroll
        for (int i = 0; i < 9; ++i) X[i] = M[i] * sc;
        // normalised determinant = s1 s2 s3 / |M|_F^3 : tiny -> ill conditioned -> SVD route
        ok = d0 * sc * sc * sc > 1e-9;
    }
    if (ok) {
        bool conv = false;
        for (int it = 0; it < 24; ++it) {
            double C[9];
            const double det = cof3(X, C);
            const double rdet = 1.0 / det;
            const double nx2 = frob2_3(X), nc2 = frob2_3(C) * rdet * rdet;
            const double g = sqrt(sqrt(nc2 / nx2));
            const double a = 0.5 * g, b = 0.5 * rdet / g;
            double diff2 = 0.0, nn2 = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double xn = a * X[i] + b * C[i];
                const double d = xn - X[i];
                diff2 += d * d; nn2 += xn * xn;
                X[i] = xn;
            }
            if (diff2 <= 1e-31 * nn2) { conv = true; break; }
        }
        ok = conv;
    }
    if (!ok) { svd3_factors(M, rot, spos, sinv); return; }
    if (rot) {
#pragma unroll
        for (int i = 0; i < 9; ++i) rot[i] = X[i];
    }
    if (spos || sinv) {
        double P[9];
        mmt3(M, X, P);   // M rot^T
        const double p01 = 0.5 * (P[1] + P[3]), p02 = 0.5 * (P[2] + P[6]), p12 = 0.5 * (P[5] + P[7]);
        P[1] = P[3] = p01; P[2] = P[6] = p02; P[5] = P[7] = p12;
        if (spos) {
#pragma unroll
            for (int i = 0; i < 9; ++i) spos[i] = P[i];
        }
        if (sinv) {
            double I[9];
            inv3(P, I);
            const double i01 = 0.5 * (I[1] + I[3]), i02 = 0.5 * (I[2] + I[6]), i12 = 0.5 * (I[5] + I[7]);
            sinv[0] = I[0]; sinv[4] = I[4]; sinv[8] = I[8];
            sinv[1] = sinv[3] = i01; sinv[2] = sinv[6] = i02; sinv[5] = sinv[7] = i12;
        }
    }
}

}  // namespace vb
