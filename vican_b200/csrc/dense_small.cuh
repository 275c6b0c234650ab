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


// ------------------------------------------------------------------------------------------
// Cooperative (warp-parallel) version of ritz9: same algorithm, every phase is a loop over
// independent work items `for (idx = lane; idx < N; idx += nl)` separated by VB_SYNC().  On the
// device one warp runs it (lane = threadIdx.x, nl = 32, arrays in shared memory) and the 9x9
// Jacobi eigen-solve uses the round-robin parallel ordering (4 disjoint rotations per round, 9
// rounds per sweep); on the host the same code runs with nl = 1.  The serial ritz9 above stays
// as the cross-check of the unit tests.
#if defined(__CUDA_ARCH__)
#define VB_SYNC() __syncwarp()
#else
#define VB_SYNC() do { } while (0)
#endif

// round-robin tournament for 9 columns (+1 dummy): round r pairs, -1 = sits out
VB_HD void rr_pair(int r, int m, int* p, int* q) {
    // circle method on 10 players, player 9 fixed (dummy), others rotate
    int a, b;
    if (m == 0) { a = 9; b = r % 9; }
    else { a = (r + m) % 9; b = (r + 9 - m) % 9; }
    if (a == 9 || b == 9) { *p = -1; *q = -1; return; }
    if (a < b) { *p = a; *q = b; } else { *p = b; *q = a; }
}

// scratch: work >= 5*81 + 64 doubles ; iwork >= 16 ints
VB_HD void ritz9_coop(const double* Gin, const double* Min, const int* act, double* C, double* Cp, double* theta,
                      int* actP, double* work, int* iwork, int lane, int nl) {
    double* G = work;            // 81
    double* M = work + 81;       // 81
    double* R = work + 162;      // 81 Cholesky factor
    double* Ri = work + 243;     // 81 its inverse
    double* Q = work + 324;      // 81 temp (G Ri), then eigenvectors
    double* sc = work + 405;     // small scratch: [0..7] c,s of 4 pairs; [8] scale/floor; [9] dot; [10..18] MZ; [19..27] z; [28..36] lam
    int* order = iwork;          // 9
    int* flags = iwork + 9;      // [0] rotations in sweep
    const int n = 9;
    for (int idx = lane; idx < 81; idx += nl) {
        const int i = idx / n, j = idx - n * i;
        const bool on = act[i] && act[j];
        double g = on ? 0.5 * (Gin[i * n + j] + Gin[j * n + i]) : 0.0;
        double m = on ? 0.5 * (Min[i * n + j] + Min[j * n + i]) : 0.0;
        if (i == j && !act[i]) { m = 1.0; g = VB_BIG; }
        G[idx] = g; M[idx] = m; R[idx] = 0.0; Ri[idx] = 0.0;
    }
    VB_SYNC();
    // Cholesky M = R^T R (upper R)
    for (int j = 0; j < n; ++j) {
        if (lane == 0) {
            double s = M[j * n + j];
            for (int k = 0; k < j; ++k) s -= R[k * n + j] * R[k * n + j];
            R[j * n + j] = sqrt(fmax(s, 1e-300));
        }
        VB_SYNC();
        for (int i = j + 1 + lane; i < n; i += nl) {
            double v = M[j * n + i];
            for (int k = 0; k < j; ++k) v -= R[k * n + j] * R[k * n + i];
            R[j * n + i] = v / R[j * n + j];
        }
        VB_SYNC();
    }
    // Ri = R^-1 (upper): one column per work item
    for (int j = lane; j < n; j += nl) {
        Ri[j * n + j] = 1.0 / R[j * n + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1; k <= j; ++k) s += R[i * n + k] * Ri[k * n + j];
            Ri[i * n + j] = -s / R[i * n + i];
        }
    }
    VB_SYNC();
    // Q = G Ri ; G = Ri^T Q ; symmetrise
    for (int idx = lane; idx < 81; idx += nl) {
        const int i = idx / n, j = idx - n * i;
        double s = 0.0;
        for (int k = 0; k <= j; ++k) s += G[i * n + k] * Ri[k * n + j];
        Q[idx] = s;
    }
    VB_SYNC();
    for (int idx = lane; idx < 81; idx += nl) {
        const int i = idx / n, j = idx - n * i;
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s += Ri[k * n + i] * Q[k * n + j];
        G[idx] = s;
    }
    VB_SYNC();
    for (int idx = lane; idx < 81; idx += nl) {
        const int i = idx / n, j = idx - n * i;
        if (i < j) { const double s = 0.5 * (G[i * n + j] + G[j * n + i]); R[i * n + j] = s; R[j * n + i] = s; }
        else if (i == j) R[idx] = G[idx];
    }
    VB_SYNC();
    for (int idx = lane; idx < 81; idx += nl) {
        const int i = idx / n, j = idx - n * i;
        G[idx] = R[idx];
        Q[idx] = (i == j) ? 1.0 : 0.0;
    }
    VB_SYNC();
    // scale for the absolute "negligible" floor (masked columns carry VB_BIG and do not count)
    if (lane == 0) {
        double scale = 0.0;
        for (int i = 0; i < 81; ++i) { const double v = fabs(G[i]); if (v < 1e-2 * VB_BIG && v > scale) scale = v; }
        sc[8] = 1e-19 * scale + 1e-300;
    }
    VB_SYNC();
    const double floor_abs = sc[8];
    // parallel-ordered cyclic Jacobi
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (lane == 0) flags[0] = 0;
        VB_SYNC();
        for (int r = 0; r < 9; ++r) {
            for (int m = lane; m < 5; m += nl) {
                int p, q;
                rr_pair(r, m, &p, &q);
                double c = 1.0, s = 0.0;
                if (p >= 0) {
                    const double apq = G[p * n + q], app = G[p * n + p], aqq = G[q * n + q];
                    if (!(fabs(apq) <= floor_abs || fabs(apq) <= 1e-17 * sqrt(fabs(app) * fabs(aqq)))) {
                        const double z = (aqq - app) / (2.0 * apq);
                        const double t = (z >= 0.0 ? 1.0 : -1.0) / (fabs(z) + sqrt(1.0 + z * z));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = c * t;
                        flags[1 + m] = 1;
                    } else flags[1 + m] = 0;
                } else flags[1 + m] = 0;
                sc[2 * m] = c; sc[2 * m + 1] = s;
            }
            VB_SYNC();
            // columns p,q of G and Q  (item = pair m, row i)
            for (int idx = lane; idx < 5 * 9; idx += nl) {
                const int m = idx / 9, i = idx - 9 * m;
                int p, q;
                rr_pair(r, m, &p, &q);
                if (p < 0) continue;
                const double c = sc[2 * m], s = sc[2 * m + 1];
                const double gp = G[i * n + p], gq = G[i * n + q];
                G[i * n + p] = c * gp - s * gq;
                G[i * n + q] = s * gp + c * gq;
                const double qp = Q[i * n + p], qq = Q[i * n + q];
                Q[i * n + p] = c * qp - s * qq;
                Q[i * n + q] = s * qp + c * qq;
            }
            VB_SYNC();
            // rows p,q of G  (item = pair m, column j)
            for (int idx = lane; idx < 5 * 9; idx += nl) {
                const int m = idx / 9, j = idx - 9 * m;
                int p, q;
                rr_pair(r, m, &p, &q);
                if (p < 0) continue;
                const double c = sc[2 * m], s = sc[2 * m + 1];
                const double gp = G[p * n + j], gq = G[q * n + j];
                G[p * n + j] = c * gp - s * gq;
                G[q * n + j] = s * gp + c * gq;
            }
            VB_SYNC();
            for (int m = lane; m < 5; m += nl) {
                int p, q;
                rr_pair(r, m, &p, &q);
                if (p >= 0) { G[p * n + q] = 0.0; G[q * n + p] = 0.0; if (flags[1 + m]) flags[0] = 1; }
            }
            VB_SYNC();
        }
        if (flags[0] == 0) break;
    }
    // sort ascending by rank counting
    double* lam = sc + 28;
    for (int i = lane; i < n; i += nl) lam[i] = G[i * n + i];
    VB_SYNC();
    for (int i = lane; i < n; i += nl) {
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (lam[j] < lam[i]) || (lam[j] == lam[i] && j < i);
        order[rank] = i;
    }
    VB_SYNC();
    for (int idx = lane; idx < 27; idx += nl) {
        const int i = idx / 3, j = idx - 3 * i;
        const int col = order[j];
        double s = 0.0;
        for (int k = i; k < n; ++k) s += Ri[i * n + k] * Q[k * n + col];
        C[idx] = s;
        if (i == 0) theta[j] = lam[col];
    }
    VB_SYNC();
    // new search directions: Z = [0; C_w; C_p], M-orthogonalised against C and the accepted
    // columns (modified Gram-Schmidt, twice), dropped when nothing is left
    double* MZ = sc + 10;
    double* z = sc + 19;
    for (int j = 0; j < 3; ++j) {
        for (int i = lane; i < n; i += nl) z[i] = (i < 3) ? 0.0 : C[i * 3 + j];
        VB_SYNC();
        double n0 = 0.0;
        for (int pass = 0; pass < 3; ++pass) {       // pass 0: norm before; 1,2: projections
            for (int cidx = 0; cidx < ((pass == 0) ? 1 : 3 + j); ++cidx) {
                for (int i = lane; i < n; i += nl) {
                    double s = 0.0;
                    for (int k = 0; k < n; ++k) s += M[i * n + k] * z[k];
                    MZ[i] = s;
                }
                VB_SYNC();
                if (pass == 0) {
                    if (lane == 0) { double s = 0.0; for (int i = 0; i < n; ++i) s += z[i] * MZ[i]; sc[9] = sqrt(fmax(s, 0.0)); }
                    VB_SYNC();
                    n0 = sc[9];
                    continue;
                }
                const double* b = (cidx < 3) ? C : Cp;
                const int bc = (cidx < 3) ? cidx : cidx - 3;
                const bool use = (cidx < 3) || actP[bc];
                if (lane == 0) { double s = 0.0; if (use) for (int i = 0; i < n; ++i) s += b[i * 3 + bc] * MZ[i]; sc[9] = s; }
                VB_SYNC();
                const double dot = sc[9];
                for (int i = lane; i < n; i += nl) z[i] -= dot * b[i * 3 + bc];
                VB_SYNC();
            }
        }
        for (int i = lane; i < n; i += nl) {
            double s = 0.0;
            for (int k = 0; k < n; ++k) s += M[i * n + k] * z[k];
            MZ[i] = s;
        }
        VB_SYNC();
        if (lane == 0) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += z[i] * MZ[i];
            const double nz = sqrt(fmax(s, 0.0));
            const bool keep = (nz > 1e-8 * fmax(n0, 1e-300)) && (nz > 1e-150);
            actP[j] = keep ? 1 : 0;
            sc[9] = keep ? 1.0 / nz : 0.0;
        }
        VB_SYNC();
        const double inz = sc[9];
        for (int i = lane; i < n; i += nl) Cp[i * 3 + j] = z[i] * inz;
        VB_SYNC();
    }
}

}  // namespace vb
