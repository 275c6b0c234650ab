// Small dense kernels of the width-3 LOBPCG (9x9 Rayleigh-Ritz, 3x3 SVQB).
//
// These replace what the reference gets from ARPACK + SuperLU inside
// scipy.sparse.linalg.eigs (vican/bipgo.py:288): the invariant subspace of the three
// eigenvalues nearest zero.  All routines are single-threaded, work on caller-provided
// arrays (shared or local memory) and are __host__ __device__ so they are unit-tested
// on the CPU against numpy (tests/test_host_math.py).
#pragma once
#include "mat3.cuh"

namespace vb {

#define VB_BIG 1e30

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9).
// A is destroyed (its diagonal holds the eigenvalues), Q gets the eigenvectors in
// columns; then eigenpairs are sorted ascending into lam / Q.
VB_HD void jacobi_eig_sym(int n, double* A, double* Q, double* lam) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Q[i * n + j] = (i == j) ? 1.0 : 0.0;
    // absolute floor for "negligible" off-diagonals: relative to the largest genuine
    // entry (masked columns carry VB_BIG on the diagonal and must not set the scale)
    double scale = 0.0;
    for (int i = 0; i < n * n; ++i) {
        const double v = fabs(A[i]);
        if (v < 1e-2 * VB_BIG && v > scale) scale = v;
    }
    const double floor_abs = 1e-19 * scale + 1e-300;
    for (int sweep = 0; sweep < 60; ++sweep) {
        int nrot = 0;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                const double app = A[p * n + p], aqq = A[q * n + q];
                if (fabs(apq) <= floor_abs || fabs(apq) <= 1e-17 * sqrt(fabs(app) * fabs(aqq))) {
                    A[p * n + q] = A[q * n + p] = 0.0;
                    continue;
                }
                ++nrot;
                const double z = (aqq - app) / (2.0 * apq);
                const double t = (z >= 0.0 ? 1.0 : -1.0) / (fabs(z) + sqrt(1.0 + z * z));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < n; ++k) {  // columns p,q
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // rows p,q
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                A[p * n + q] = A[q * n + p] = 0.0;
                for (int k = 0; k < n; ++k) {
                    const double qkp = Q[k * n + p], qkq = Q[k * n + q];
                    Q[k * n + p] = c * qkp - s * qkq;
                    Q[k * n + q] = s * qkp + c * qkq;
                }
            }
        if (nrot == 0) break;
    }
    for (int i = 0; i < n; ++i) lam[i] = A[i * n + i];
    // selection sort ascending
    for (int i = 0; i < n - 1; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j)
            if (lam[j] < lam[m]) m = j;
        if (m != i) {
            const double t = lam[i]; lam[i] = lam[m]; lam[m] = t;
            for (int k = 0; k < n; ++k) { const double u = Q[k * n + i]; Q[k * n + i] = Q[k * n + m]; Q[k * n + m] = u; }
        }
    }
}

// SVQB orthonormalisation map for a 3-column block with Gram matrix Gw (3x3):
// returns T (3x3) such that (W T) has orthonormal columns; directions whose scaled
// Gram eigenvalue is below drop_tol * max are dropped (T column = 0, act = 0).
VB_HD void svqb3(const double* Gw, double* T, int* act, double drop_tol) {
    double d[3], Gs[9], Q[9], lam[3];
    for (int i = 0; i < 3; ++i) d[i] = sqrt(fmax(Gw[4 * i], 1e-300));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Gs[3 * i + j] = 0.5 * (Gw[3 * i + j] + Gw[3 * j + i]) / (d[i] * d[j]);
    jacobi_eig_sym(3, Gs, Q, lam);
    const double lmax = fmax(lam[2], 1e-300);
    for (int j = 0; j < 3; ++j) {
        const bool keep = lam[j] > drop_tol * lmax;
        act[j] = keep ? 1 : 0;
        const double sc = keep ? 1.0 / sqrt(lam[j]) : 0.0;
        for (int i = 0; i < 3; ++i) T[3 * i + j] = Q[3 * i + j] / d[i] * sc;
    }
}

// Rayleigh-Ritz on the basis S = [X W P] (9 columns).
//   in : G = S^T A S (9x9), M = S^T S (9x9), act[9] (1 = column present)
//   out: C  (9x3)  coefficients of the new X       (X' = S C, M-orthonormal)
//        Cp (9x3)  coefficients of the new P       (P' = S Cp, M-orthonormal, M-orth. to X')
//        theta[3]  the three smallest Ritz values, actP[3]
// scratch: work >= 4*81 doubles.
VB_HD void ritz9(const double* Gin, const double* Min, const int* act, double* C, double* Cp, double* theta,
                 int* actP, double* work) {
    double* G = work;          // 81
    double* M = work + 81;     // 81
    double* R = work + 162;    // 81 upper Cholesky factor, then its inverse
    double* Q = work + 243;    // 81
    const int n = 9;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const bool on = act[i] && act[j];
            G[i * n + j] = on ? 0.5 * (Gin[i * n + j] + Gin[j * n + i]) : 0.0;
            M[i * n + j] = on ? 0.5 * (Min[i * n + j] + Min[j * n + i]) : 0.0;
        }
    for (int j = 0; j < n; ++j)
        if (!act[j]) { M[j * n + j] = 1.0; G[j * n + j] = VB_BIG; }
    // Cholesky M = R^T R (upper R)
    for (int i = 0; i < n * n; ++i) R[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double s = M[j * n + j];
        for (int k = 0; k < j; ++k) s -= R[k * n + j] * R[k * n + j];
        const double rjj = sqrt(fmax(s, 1e-300));
        R[j * n + j] = rjj;
        for (int i = j + 1; i < n; ++i) {
            double v = M[j * n + i];
            for (int k = 0; k < j; ++k) v -= R[k * n + j] * R[k * n + i];
            R[j * n + i] = v / rjj;
        }
    }
    // Ri = R^-1 (upper), stored into Q temporarily then copied to R
    for (int i = 0; i < n * n; ++i) Q[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        Q[j * n + j] = 1.0 / R[j * n + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1; k <= j; ++k) s += R[i * n + k] * Q[k * n + j];
            Q[i * n + j] = -s / R[i * n + i];
        }
    }
    for (int i = 0; i < n * n; ++i) R[i] = Q[i];   // R now holds Ri
    // Gw = Ri^T G Ri   (into M-scratch: keep M intact -> use Q as temp, result in G)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k <= j; ++k) s += G[i * n + k] * R[k * n + j];
            Q[i * n + j] = s;   // G Ri
        }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 0; k <= i; ++k) s += R[k * n + i] * Q[k * n + j];
            G[i * n + j] = s;   // Ri^T (G Ri)
        }
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) { const double s = 0.5 * (G[i * n + j] + G[j * n + i]); G[i * n + j] = G[j * n + i] = s; }
    double lam[9];
    jacobi_eig_sym(n, G, Q, lam);
    for (int j = 0; j < 3; ++j) {
        theta[j] = lam[j];
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int k = i; k < n; ++k) s += R[i * n + k] * Q[k * n + j];
            C[i * 3 + j] = s;
        }
    }
    // new search directions: Z = [0; C_w; C_p], M-orthogonalised against C, then
    // M-orthonormalised by modified Gram-Schmidt (twice) with dropping.
    double Z[27], MZ[9];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) Z[i * 3 + j] = (i < 3) ? 0.0 : C[i * 3 + j];
    for (int j = 0; j < 3; ++j) {
        // MZ = M z_j
        for (int i = 0; i < n; ++i) { double s = 0.0; for (int k = 0; k < n; ++k) s += M[i * n + k] * Z[k * 3 + j]; MZ[i] = s; }
        double n0 = 0.0;
        for (int i = 0; i < n; ++i) n0 += Z[i * 3 + j] * MZ[i];
        n0 = sqrt(fmax(n0, 0.0));
        for (int pass = 0; pass < 2; ++pass) {
            for (int c = 0; c < 3 + j; ++c) {
                const double* b = (c < 3) ? C : Cp;
                const int bc = (c < 3) ? c : c - 3;
                if (c >= 3 && !actP[bc]) continue;
                for (int i = 0; i < n; ++i) { double s = 0.0; for (int k = 0; k < n; ++k) s += M[i * n + k] * Z[k * 3 + j]; MZ[i] = s; }
                double dot = 0.0;
                for (int i = 0; i < n; ++i) dot += b[i * 3 + bc] * MZ[i];
                for (int i = 0; i < n; ++i) Z[i * 3 + j] -= dot * b[i * 3 + bc];
            }
        }
        for (int i = 0; i < n; ++i) { double s = 0.0; for (int k = 0; k < n; ++k) s += M[i * n + k] * Z[k * 3 + j]; MZ[i] = s; }
        double nz = 0.0;
        for (int i = 0; i < n; ++i) nz += Z[i * 3 + j] * MZ[i];
        nz = sqrt(fmax(nz, 0.0));
        const bool keep = (nz > 1e-8 * fmax(n0, 1e-300)) && (nz > 1e-150);
        actP[j] = keep ? 1 : 0;
        for (int i = 0; i < n; ++i) Cp[i * 3 + j] = keep ? Z[i * 3 + j] / nz : 0.0;
    }
}

}  // namespace vb
