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

// 1/sqrt(x): the device intrinsic (MUFU.RSQ64H + Newton, ~1 ulp) is several times cheaper than a
// sqrt followed by a division, and the Jacobi / Cholesky chains below are pure latency
VB_HD double vb_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// Jacobi rotation (c, s) annihilating a_pq, the smaller of the two roots (|tan| <= 1): with
// d = a_qq - a_pp, w = 2 a_pq, r = sqrt(d^2 + w^2):  cos(2 theta) = |d| / r,
//   c = sqrt((1 + cos 2theta) / 2),   s = sgn(d) w / (2 r c)      (= c tan, tan = sgn(d) w / (|d| + r)).
// Two dependent reciprocal square roots and no division: the chain of the serial part of a round.
VB_HD void jacobi_cs(double app, double aqq, double apq, double* c, double* s) {
    const double d = aqq - app, w = 2.0 * apq;
    const double inv_r = vb_rsqrt(d * d + w * w);
    const double x = 0.5 + 0.5 * (fabs(d) * inv_r);      // in [0.5, 1]
    const double rs = vb_rsqrt(x);
    *c = x * rs;
    *s = (d >= 0.0 ? w : -w) * inv_r * (0.5 * rs);
}
// "negligible" test without a square root: |a_pq| <= 1e-17 sqrt(|a_pp a_qq|)
VB_HD bool jacobi_negligible(double app, double aqq, double apq, double floor_abs) {
    return fabs(apq) <= floor_abs || apq * apq <= 1e-34 * fabs(app * aqq);
}

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
                if (jacobi_negligible(app, aqq, apq, floor_abs)) {
                    A[p * n + q] = A[q * n + p] = 0.0;
                    continue;
                }
                ++nrot;
                double c, s;
                jacobi_cs(app, aqq, apq, &c, &s);
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
    // (the register-resident one-sided Jacobi SVD of mat3.cuh was tried here: it needs many more sweeps on a
    // Gram matrix with a 1e-12 eigenvalue spread and measured 2-4x slower than the generic two-sided Jacobi)
    double d[3], Gs[9], Q[9], lam[3];
    for (int i = 0; i < 3; ++i) d[i] = vb_rsqrt(fmax(Gw[4 * i], 1e-300));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Gs[3 * i + j] = 0.5 * (Gw[3 * i + j] + Gw[3 * j + i]) * (d[i] * d[j]);
    jacobi_eig_sym(3, Gs, Q, lam);
    const double lmax = fmax(lam[2], 1e-300);
    for (int j = 0; j < 3; ++j) {
        const bool keep = lam[j] > drop_tol * lmax;
        act[j] = keep ? 1 : 0;
        const double sc = keep ? vb_rsqrt(lam[j]) : 0.0;
        for (int i = 0; i < 3; ++i) T[3 * i + j] = Q[3 * i + j] * d[i] * sc;
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

// round-robin tournament (circle method) on np players, np even: in round r (0 <= r < np-1) pair m
// (0 <= m < np/2) is (a, b); with an odd number of real columns the last player is a dummy and its
// pair sits out (*p = -1)
VB_HD void rr_pair(int np, int n_real, int r, int m, int* p, int* q) {
    const int k = np - 1;
    int a, b;
    if (m == 0) { a = k; b = r % k; }
    else { a = (r + m) % k; b = (r + k - m) % k; }
    if (a >= n_real || b >= n_real) { *p = -1; *q = -1; return; }
    if (a < b) { *p = a; *q = b; } else { *p = b; *q = a; }
}

// Parallel-ordered cyclic Jacobi sweeps on a compact NA x NA symmetric matrix (row stride 9) with
// eigenvector accumulation, double buffered: Ga/Qa hold the input, the result is in *Gout / *Qout.
// One round = 2 phases: (A) rotation parameters of the disjoint pairs -> per-column tables
// (partner, c, s); (B) G' = J^T G J and Q' = Q J written from the old matrices into the other
// buffers.  flags[3]: "rotation seen in sweep s" in slot s % 3.  Returns the sweeps used.
template <int NA>
VB_HD int jacobi_rounds(double* Ga, double* Gb, double* Qa, double* Qb, double* sc, int* part, int* flags,
                        double floor_abs, int lane, int nl, double** Gout, double** Qout) {
    constexpr int n = 9;
    constexpr int NP = (NA + 1) & ~1;          // players of the tournament (dummy added when NA is odd)
    constexpr int ROUNDS = NP - 1, PAIRS = NP / 2;
    double *Gc = Ga, *Gn = Gb, *Qc = Qa, *Qn = Qb;
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        for (int r = 0; r < ROUNDS; ++r) {
            for (int m = lane; m < PAIRS; m += nl) {
                int p, q;
                rr_pair(NP, NA, r, m, &p, &q);
                if (m == 0) {
                    if (r == 0) flags[(sweep + 1) % 3] = 0;               // next sweep's slot: last read two sweeps ago
                    if (NA & 1) { const int lone = r % ROUNDS; part[lone] = lone; sc[lone] = 1.0; sc[9 + lone] = 0.0; }
                }
                if (p < 0) continue;
                double c = 1.0, sn = 0.0;
                const double apq = Gc[p * n + q], app = Gc[p * n + p], aqq = Gc[q * n + q];
                if (!jacobi_negligible(app, aqq, apq, floor_abs)) {
                    jacobi_cs(app, aqq, apq, &c, &sn);
                    flags[sweep % 3] = 1;
                }
                // new column p = c col_p - s col_q ; new column q = s col_p + c col_q
                part[p] = q; sc[p] = c; sc[9 + p] = -sn;
                part[q] = p; sc[q] = c; sc[9 + q] = sn;
            }
            VB_SYNC();
            for (int idx = lane; idx < NA * NA; idx += nl) {
                const int i = idx / NA, j = idx - NA * i;
                const int pi = part[i], pj = part[j];
                const double ci = sc[i], si = sc[9 + i], cj = sc[j], sj = sc[9 + j];
                const double g = ci * (cj * Gc[i * n + j] + sj * Gc[i * n + pj]) + si * (cj * Gc[pi * n + j] + sj * Gc[pi * n + pj]);
                Gn[i * n + j] = (pi == j && i != j) ? 0.0 : g;          // the annihilated pair entries
                Qn[i * n + j] = cj * Qc[i * n + j] + sj * Qc[i * n + pj];
            }
            VB_SYNC();
            double* tg = Gc; Gc = Gn; Gn = tg;
            double* tq = Qc; Qc = Qn; Qn = tq;
        }
        if (flags[sweep % 3] == 0) { ++sweep; break; }
    }
    *Gout = Gc; *Qout = Qc;
    return sweep;
}

// scratch: work >= 5*81 + 64 doubles ; iwork >= 48 ints
//
// Phases are loops over independent work items separated by VB_SYNC(); the design goal on the
// device is a SHORT CHAIN of phases (the step kernel waits on this routine with the whole grid):
//  * the problem is compacted to the active columns (3, 6 or 9 of them): 3 / 5 / 9 Jacobi rounds
//    per sweep instead of always 9;
//  * one Jacobi round = 2 phases: rotation parameters of the disjoint pairs, then the two-sided
//    update G <- J^T G J and Q <- Q J written from the OLD matrices into a second buffer (no
//    column-phase / row-phase / clean-up split);
//  * the new search directions are M-orthonormalised by classical Gram-Schmidt applied twice with
//    the products M [C Cp] cached, all dot products of a pass evaluated redundantly by every lane
//    (7 phases per column instead of ~36).
VB_HD void ritz9_coop(const double* Gin, const double* Min, const int* act, double* C, double* Cp, double* theta,
                      int* actP, double* work, int* iwork, int lane, int nl) {
    double* Ga = work;           // 81  G ping
    double* Gb = work + 81;      // 81  G pong
    double* Qa = work + 162;     // 81  Cholesky factor R, then Q ping
    double* Ri = work + 243;     // 81  R^-1
    double* Qb = work + 324;     // 81  temp (G Ri), partial maxima, then Q pong
    double* sc = work + 405;     // 64: [0..8] ci, [9..17] si, [18..26] 1/R_jj, [27..35] order, [36..44] lam
    int* a = iwork;              // [0..8]   active column list
    int* part = iwork + 9;       // [9..17]  partner of a column in the current round
    int* flags = iwork + 18;     // [18..20] "rotation seen in sweep s" in slot s % 3 (reset two sweeps ahead of its reuse)
    int* pos = iwork + 21;       // [21..29] column -> compact position (-1 = inactive)
    const int n = 9;
    // ---- compact list of the active columns (every lane computes the same small table)
    int na = 0;
    for (int i = 0; i < n; ++i) na += act[i] ? 1 : 0;
    if (lane == 0) {
        int k = 0;
        for (int i = 0; i < n; ++i) { pos[i] = act[i] ? k : -1; if (act[i]) a[k++] = i; }
        flags[0] = 0; flags[1] = 0; flags[2] = 0;
    }
    VB_SYNC();
    // compact symmetric copies: Ga = G, Gb = M (temporarily)
    for (int idx = lane; idx < na * na; idx += nl) {
        const int i = idx / na, j = idx - na * i;
        const int gi = a[i], gj = a[j];
        Ga[i * n + j] = 0.5 * (Gin[gi * n + gj] + Gin[gj * n + gi]);
        Gb[i * n + j] = 0.5 * (Min[gi * n + gj] + Min[gj * n + gi]);
        Qa[i * n + j] = 0.0; Ri[i * n + j] = 0.0;
    }
    VB_SYNC();
    // ---- Cholesky M = R^T R (upper R in Qa), reciprocal pivots in sc[18..]
    for (int j = 0; j < na; ++j) {
        if (lane == 0) {
            double s = Gb[j * n + j];
            for (int k = 0; k < j; ++k) s -= Qa[k * n + j] * Qa[k * n + j];
            s = fmax(s, 1e-300);
            const double ri = vb_rsqrt(s);
            Qa[j * n + j] = s * ri;
            sc[18 + j] = ri;
        }
        VB_SYNC();
        for (int i = j + 1 + lane; i < na; i += nl) {
            double v = Gb[j * n + i];
            for (int k = 0; k < j; ++k) v -= Qa[k * n + j] * Qa[k * n + i];
            Qa[j * n + i] = v * sc[18 + j];
        }
        VB_SYNC();
    }
    // Ri = R^-1 (upper): one column per work item
    for (int j = lane; j < na; j += nl) {
        Ri[j * n + j] = sc[18 + j];
        for (int i = j - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1; k <= j; ++k) s += Qa[i * n + k] * Ri[k * n + j];
            Ri[i * n + j] = -s * sc[18 + i];
        }
    }
    VB_SYNC();
    // ---- whitened matrix A = Ri^T G Ri: Qb = G Ri, Gb = Ri^T Qb, then symmetrise into Ga, Qa = I
    for (int idx = lane; idx < na * na; idx += nl) {
        const int i = idx / na, j = idx - na * i;
        double s = 0.0;
        for (int k = 0; k <= j; ++k) s += Ga[i * n + k] * Ri[k * n + j];
        Qb[i * n + j] = s;
    }
    VB_SYNC();
    for (int idx = lane; idx < na * na; idx += nl) {
        const int i = idx / na, j = idx - na * i;
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s += Ri[k * n + i] * Qb[k * n + j];
        Gb[i * n + j] = s;
    }
    VB_SYNC();
    {
        double mx = 0.0;
        for (int idx = lane; idx < na * na; idx += nl) {
            const int i = idx / na, j = idx - na * i;
            const double v = 0.5 * (Gb[i * n + j] + Gb[j * n + i]);
            Ga[i * n + j] = v;
            Qa[i * n + j] = (i == j) ? 1.0 : 0.0;
            mx = fmax(mx, fabs(v));
        }
        Qb[81 - 32 + (lane & 31)] = mx;     // tail of Qb: its compact matrix part is not live here
    }
    VB_SYNC();
    double scale = 0.0;
    for (int l = 0; l < nl && l < 32; ++l) scale = fmax(scale, Qb[81 - 32 + l]);
    const double floor_abs = 1e-19 * scale + 1e-300;
    // ---- parallel-ordered cyclic Jacobi on the na x na matrix (size-specialised: the index arithmetic of
    // the rounds is compile-time, a run-time `idx / na` per work item doubled the cost of a round)
    double *Gc = Ga, *Qc = Qa;
    int sweeps_used = 0;
    switch (na) {
        case 3: sweeps_used = jacobi_rounds<3>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 6: sweeps_used = jacobi_rounds<6>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 9: sweeps_used = jacobi_rounds<9>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 1: break;
        case 2: sweeps_used = jacobi_rounds<2>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 4: sweeps_used = jacobi_rounds<4>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 5: sweeps_used = jacobi_rounds<5>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        case 7: sweeps_used = jacobi_rounds<7>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
        default: sweeps_used = jacobi_rounds<8>(Ga, Gb, Qa, Qb, sc, part, flags, floor_abs, lane, nl, &Gc, &Qc); break;
    }
    if (lane == 0) iwork[30] = sweeps_used;        // diagnostics
    // ---- eigenvalues ascending (rank counting), the 3 lowest Ritz vectors in the original basis
    for (int i = lane; i < na; i += nl) {
        const double li = Gc[i * n + i];
        int rank = 0;
        for (int j = 0; j < na; ++j) { const double lj = Gc[j * n + j]; rank += (lj < li) || (lj == li && j < i); }
        sc[27 + rank] = (double)i;
        sc[36 + i] = li;
    }
    VB_SYNC();
    for (int idx = lane; idx < 27; idx += nl) {
        const int gi = idx / 3, j = idx - 3 * gi;
        const int i = pos[gi];
        double s = 0.0;
        if (i >= 0 && j < na) {
            const int col = (int)sc[27 + j];
            for (int k = i; k < na; ++k) s += Ri[i * n + k] * Qc[k * n + col];
            if (i == 0) theta[j] = sc[36 + col];
        }
        C[idx] = s;
        Cp[idx] = 0.0;
    }
    if (lane == 0) for (int j = na; j < 3; ++j) theta[j] = VB_BIG;
    VB_SYNC();
    // ---- new search directions: Z = [0; C_w; C_p] M-orthonormalised against C and the accepted
    // columns by classical Gram-Schmidt applied twice, dropped when nothing is left.
    // MB[:, c] = M B[:, c] for B = [C | Cp] is cached (Ga: 9 x 6), z in Gb[0..8], M z in Gb[9..17].
    double* MB = Ga;
    double* z = Gb;
    double* Mz = Gb + 9;
    auto Msym = [&](int i, int k) { return 0.5 * (Min[i * n + k] + Min[k * n + i]); };
    for (int idx = lane; idx < 27; idx += nl) {
        const int i = idx / 3, c = idx - 3 * i;
        double s = 0.0;
        if (act[i]) for (int k = 0; k < n; ++k) if (act[k]) s += Msym(i, k) * C[k * 3 + c];
        MB[i * 6 + c] = s;
    }
    for (int j = 0; j < 3; ++j) {
        for (int i = lane; i < n; i += nl) z[i] = (i < 3) ? 0.0 : C[i * 3 + j];
        VB_SYNC();
        for (int i = lane; i < n; i += nl) {
            double s = 0.0;
            if (act[i]) for (int k = 0; k < n; ++k) if (act[k]) s += Msym(i, k) * z[k];
            Mz[i] = s;
        }
        VB_SYNC();
        double n0 = 0.0;
        for (int i = 0; i < n; ++i) n0 += z[i] * Mz[i];
        n0 = sqrt(fmax(n0, 0.0));
        for (int pass = 0; pass < 2; ++pass) {
            double d[5];
            for (int c = 0; c < 3 + j; ++c) {
                double s = 0.0;
                if (c < 3 || actP[c - 3]) for (int i = 0; i < n; ++i) s += MB[i * 6 + c] * z[i];
                d[c] = s;
            }
            VB_SYNC();                       // every lane has read z before anybody updates it
            for (int i = lane; i < n; i += nl) {
                double v = z[i];
                for (int c = 0; c < 3 + j; ++c) v -= d[c] * ((c < 3) ? C[i * 3 + c] : Cp[i * 3 + (c - 3)]);
                z[i] = v;
            }
            VB_SYNC();
        }
        for (int i = lane; i < n; i += nl) {
            double s = 0.0;
            if (act[i]) for (int k = 0; k < n; ++k) if (act[k]) s += Msym(i, k) * z[k];
            Mz[i] = s;
        }
        VB_SYNC();
        double nz = 0.0;
        for (int i = 0; i < n; ++i) nz += z[i] * Mz[i];
        nz = sqrt(fmax(nz, 0.0));
        const bool keep = (nz > 1e-8 * fmax(n0, 1e-300)) && (nz > 1e-150);
        const double inz = keep ? 1.0 / nz : 0.0;
        for (int i = lane; i < n; i += nl) {
            Cp[i * 3 + j] = z[i] * inz;
            MB[i * 6 + 3 + j] = Mz[i] * inz;
        }
        if (lane == 0) actP[j] = keep ? 1 : 0;
        VB_SYNC();
    }
}

}  // namespace vb
