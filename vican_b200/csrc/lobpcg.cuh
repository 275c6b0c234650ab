// Width-3 LOBPCG step for the 3 eigenpairs of L = Lambda_C - P Lambda_T P^T nearest zero.
//
// Replaces scipy.sparse.linalg.eigs(L, k=5, sigma=-1e-6) (ARPACK shift-invert + SuperLU,
// vican/bipgo.py:288): only the 3-dimensional invariant subspace enters the result
// (bipgo.py:295), and it is independent of the eigensolver.
//
// One cooperative kernel does ALL vector work between two applications of L:
//   stage 1  A-image of the freshly applied block (AW = Lambda_C W - Y), Gram matrices
//            G = S^T A S and M = S^T S of the basis S = [X W P]                  (grid reduce 1)
//   stage 2  9x9 Rayleigh-Ritz, redundantly per CTA (dense_small.cuh: ritz9)
//   stage 3  X,AX,P,AP <- basis update; residual R = AX - X theta; W = Lambda_C^-1 R
//            (block-Jacobi preconditioner); H = [X P]^T W, ||R_j||^2           (grid reduce 2)
//            -> convergence test (uniform across CTAs)
//   stage 4  W -= [X P] H; second projection coefficients and Gram of W         (grid reduce 3)
//   stage 5  W -= [X P] H2; W <- W T  (SVQB orthonormalisation with dropping)
// Block vectors are stored as [n_c][9] (camera block, row-major 3x3): row (c,i), column j.
// Camera-side vectors are tiny (72 B per camera per vector), so this kernel is latency- not
// bandwidth-bound; it exists to keep the whole eigen-iteration on the device.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "dense_small.cuh"

namespace vb {
namespace cg = cooperative_groups;

constexpr int LOB_THREADS = 256;
constexpr int LOB_NRED = 112;      // >= 12*9 + 1
// persistent small state (doubles)
constexpr int SM_THETA = 0, SM_RESN = 3, SM_CONV = 6, SM_ANORM = 7, SM_ITERS = 8, SM_ACT = 9, SM_TIME = 18, SM_SIZE = 32;
// SM_TIME..+9: globaltimer stamps (ns) of block 0 at the stage boundaries of the last step (diagnostics)

struct LobpcgParams {
    int n_c;
    double *X, *AX, *W, *AW, *P, *AP;
    const double* Y;
    const double *lamC, *lamCinv;
    double* small;
    double* partial;   // [3][gridDim.x][LOB_NRED]
    double tol;
    int first;
};

template <int N>
__device__ __forceinline__ void block_reduce_store(double (&v)[N], double* sm /*[8][N]*/, double* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) sm[warp * N + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < LOB_THREADS / 32; ++w) s += sm[w * N + threadIdx.x];
        dst[threadIdx.x] = s;
    }
    __syncthreads();
}

// deterministic: every CTA sums the per-CTA partials in the same order
__device__ __forceinline__ void grid_combine(const double* partial, int n, double* out_sm) {
    if ((int)threadIdx.x < n) {
        double s = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) s += partial[(size_t)b * LOB_NRED + threadIdx.x];
        out_sm[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void ld9(const double* p, double* v) {
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = p[i];
}
__device__ __forceinline__ void st9(double* p, const double* v) {
#pragma unroll
    for (int i = 0; i < 9; ++i) p[i] = v[i];
}
// acc[3j+jp] += sum_i a[3i+j] b[3i+jp]      (a^T b for row-major 3x3 blocks)
__device__ __forceinline__ void atb_acc(const double* a, const double* b, double* acc) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int jp = 0; jp < 3; ++jp)
            acc[3 * j + jp] += a[j] * b[jp] + a[3 + j] * b[3 + jp] + a[6 + j] * b[6 + jp];
}
// out[3i+j] (+)= sum_jj v[3i+jj] * Cm[(jj)*ldc + j]
__device__ __forceinline__ void blk_times(const double* v, const double* Cm, int ldc, double* out) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            out[3 * i + j] += v[3 * i] * Cm[j] + v[3 * i + 1] * Cm[ldc + j] + v[3 * i + 2] * Cm[2 * ldc + j];
}

__device__ __forceinline__ void lob_stamp(const LobpcgParams& p, int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.small[SM_TIME + k] = (double)t;
    }
}

__global__ void __launch_bounds__(LOB_THREADS, 1) lobpcg_step_kernel(LobpcgParams p) {
    // a speculatively enqueued step after convergence is a no-op (uniform across the grid: the flag
    // was written by the previous launch)
    if (!p.first && p.small[SM_CONV] != 0.0) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ double red_sm[8 * 27];
    __shared__ double tot[LOB_NRED];
    __shared__ double Gm[81], Mm[81], Cx[27], Cp[27], work[5 * 81 + 64];
    __shared__ int iwork[48];
    __shared__ double theta_s[3], H[18], T[9];
    __shared__ int act_s[9], actP_s[3], actW_s[3], conv_s;

    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gthreads = gridDim.x * blockDim.x;
    const int n_c = p.n_c;
    const double* base1 = p.partial;
    const double* base2 = p.partial + (size_t)gridDim.x * LOB_NRED;
    const double* base3 = p.partial + (size_t)2 * gridDim.x * LOB_NRED;
    double* part1 = p.partial + (size_t)blockIdx.x * LOB_NRED;
    double* part2 = p.partial + ((size_t)gridDim.x + blockIdx.x) * LOB_NRED;
    double* part3 = p.partial + ((size_t)2 * gridDim.x + blockIdx.x) * LOB_NRED;

    lob_stamp(p, 0);
    // ---------------- stage 1: A-image of the applied block, Gram matrices -----------------
    {
        double* V = p.first ? p.X : p.W;
        double* AV = p.first ? p.AX : p.AW;
        double an[1] = {0.0};
        for (int c = gtid; c < n_c; c += gthreads) {
            double v[9], y[9], l[9], o[9];
            ld9(V + 9 * (size_t)c, v);
            ld9(p.Y + 9 * (size_t)c, y);
            ld9(p.lamC + 9 * (size_t)c, l);
            mm3(l, v, o);
#pragma unroll
            for (int i = 0; i < 9; ++i) { o[i] -= y[i]; an[0] += l[i] * l[i]; }
            st9(AV + 9 * (size_t)c, o);
        }
        block_reduce_store<1>(an, red_sm, part1 + 108);
        // stores above are re-read below by the same thread only (same camera stride) -> no sync needed
        const double* Vs[3] = {p.X, p.W, p.P};
        const double* AVs[3] = {p.AX, p.AW, p.AP};
        int blk = 0;
        for (int a = 0; a < 3; ++a)
            for (int b = a; b < 3; ++b, ++blk) {
                double g[9], m[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) { g[i] = 0.0; m[i] = 0.0; }
                for (int c = gtid; c < n_c; c += gthreads) {
                    double va[9], vb_[9], avb[9];
                    ld9(Vs[a] + 9 * (size_t)c, va);
                    ld9(Vs[b] + 9 * (size_t)c, vb_);
                    ld9(AVs[b] + 9 * (size_t)c, avb);
                    atb_acc(va, avb, g);
                    atb_acc(va, vb_, m);
                }
                block_reduce_store<9>(g, red_sm, part1 + 9 * blk);
                block_reduce_store<9>(m, red_sm, part1 + 54 + 9 * blk);
            }
    }
    lob_stamp(p, 1);
    grid.sync();
    lob_stamp(p, 2);
    grid_combine(base1, 109, tot);
    lob_stamp(p, 3);

    // ---------------- stage 2: Rayleigh-Ritz (redundant per CTA, deterministic) ------------
    // one warp per CTA: warp-cooperative 9x9 solve (parallel-ordered Jacobi), dense_small.cuh
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int idx = lane; idx < 54; idx += 32) {
            const int blk = idx / 9, e9 = idx - 9 * blk, j = e9 / 3, jp = e9 - 3 * j;
            // upper blocks in the order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
            const int a = (blk < 3) ? 0 : ((blk < 5) ? 1 : 2);
            const int b = (blk < 3) ? blk : ((blk < 5) ? blk - 2 : 2);
            const double g = tot[idx], m = tot[54 + idx];
            Gm[(3 * a + j) * 9 + 3 * b + jp] = g;
            Mm[(3 * a + j) * 9 + 3 * b + jp] = m;
            if (a != b) {
                Gm[(3 * b + jp) * 9 + 3 * a + j] = g;
                Mm[(3 * b + jp) * 9 + 3 * a + j] = m;
            }
        }
        for (int i = lane; i < 9; i += 32) act_s[i] = (i < 3) ? 1 : (p.first ? 0 : (p.small[SM_ACT + i] != 0.0));
        __syncwarp();
        ritz9_coop(Gm, Mm, act_s, Cx, Cp, theta_s, actP_s, work, iwork, lane, 32);
    }
    __syncthreads();
    lob_stamp(p, 4);

    // ---------------- stage 3: basis update, residual, preconditioned direction ------------
    const double anorm = sqrt(tot[108] / (3.0 * n_c));
    {
        double acc[21];
#pragma unroll
        for (int i = 0; i < 21; ++i) acc[i] = 0.0;
        for (int c = gtid; c < n_c; c += gthreads) {
            const size_t o = 9 * (size_t)c;
            double x[9], w[9], pp[9], xn[9], pn[9];
            ld9(p.X + o, x); ld9(p.W + o, w); ld9(p.P + o, pp);
#pragma unroll
            for (int i = 0; i < 9; ++i) { xn[i] = 0.0; pn[i] = 0.0; }
            blk_times(x, Cx, 3, xn);      blk_times(w, Cx + 9, 3, xn);  blk_times(pp, Cx + 18, 3, xn);
            blk_times(x, Cp, 3, pn);      blk_times(w, Cp + 9, 3, pn);  blk_times(pp, Cp + 18, 3, pn);
            st9(p.X + o, xn); st9(p.P + o, pn);
            double ax[9], aw[9], ap[9], axn[9], apn[9];
            ld9(p.AX + o, ax); ld9(p.AW + o, aw); ld9(p.AP + o, ap);
#pragma unroll
            for (int i = 0; i < 9; ++i) { axn[i] = 0.0; apn[i] = 0.0; }
            blk_times(ax, Cx, 3, axn);    blk_times(aw, Cx + 9, 3, axn); blk_times(ap, Cx + 18, 3, axn);
            blk_times(ax, Cp, 3, apn);    blk_times(aw, Cp + 9, 3, apn); blk_times(ap, Cp + 18, 3, apn);
            st9(p.AX + o, axn); st9(p.AP + o, apn);
            double r[9], li[9], wn[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    r[3 * i + j] = axn[3 * i + j] - xn[3 * i + j] * theta_s[j];
                    acc[18 + j] += r[3 * i + j] * r[3 * i + j];
                }
            ld9(p.lamCinv + o, li);
            mm3(li, r, wn);
            st9(p.W + o, wn);
            atb_acc(xn, wn, acc);
            atb_acc(pn, wn, acc + 9);
        }
        block_reduce_store<21>(acc, red_sm, part2);
    }
    lob_stamp(p, 5);
    grid.sync();
    grid_combine(base2, 21, tot);
    lob_stamp(p, 6);
    if (threadIdx.x == 0) {
        const double r0 = sqrt(tot[18]), r1 = sqrt(tot[19]), r2 = sqrt(tot[20]);
        const double rmax = fmax(r0, fmax(r1, r2));
        conv_s = (rmax <= p.tol * anorm) ? 1 : 0;
        for (int i = 0; i < 18; ++i) H[i] = tot[i];
        if (blockIdx.x == 0) {
            p.small[SM_THETA] = theta_s[0]; p.small[SM_THETA + 1] = theta_s[1]; p.small[SM_THETA + 2] = theta_s[2];
            p.small[SM_RESN] = r0; p.small[SM_RESN + 1] = r1; p.small[SM_RESN + 2] = r2;
            p.small[SM_CONV] = (double)conv_s;
            p.small[SM_ANORM] = anorm;
            p.small[SM_ITERS] = p.first ? 1.0 : p.small[SM_ITERS] + 1.0;
            for (int j = 0; j < 3; ++j) p.small[SM_ACT + 6 + j] = (double)actP_s[j];
        }
    }
    __syncthreads();
    if (conv_s) return;   // uniform over the grid: every CTA evaluated the same numbers

    // ---------------- stage 4: project W against [X P], Gram of W ---------------------------
    {
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0.0;
        for (int c = gtid; c < n_c; c += gthreads) {
            const size_t o = 9 * (size_t)c;
            double x[9], pp[9], w[9];
            ld9(p.X + o, x); ld9(p.P + o, pp); ld9(p.W + o, w);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double s = 0.0;
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) s += x[3 * i + jj] * H[3 * jj + j] + pp[3 * i + jj] * H[9 + 3 * jj + j];
                    w[3 * i + j] -= s;
                }
            st9(p.W + o, w);
            atb_acc(x, w, acc);
            atb_acc(pp, w, acc + 9);
            atb_acc(w, w, acc + 18);
        }
        block_reduce_store<27>(acc, red_sm, part3);
    }
    lob_stamp(p, 7);
    grid.sync();
    grid_combine(base3, 27, tot);
    lob_stamp(p, 8);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 18; ++i) H[i] = tot[i];
        // Gram of the twice-projected block: the second correction is O(eps) so
        // W2^T W2 = Gw - H2^T H2 to O(eps^2); keep it symmetric.
        double Gw[9];
        for (int j = 0; j < 3; ++j)
            for (int jp = 0; jp < 3; ++jp) {
                double s = tot[18 + 3 * j + jp];
                for (int k = 0; k < 6; ++k) s -= tot[3 * k + j] * tot[3 * k + jp];
                Gw[3 * j + jp] = s;
            }
        svqb3(Gw, T, actW_s, 1e-12);
        if (blockIdx.x == 0)
            for (int j = 0; j < 3; ++j) p.small[SM_ACT + 3 + j] = (double)actW_s[j];
    }
    __syncthreads();

    // ---------------- stage 5: second projection + orthonormalisation ------------------------
    for (int c = gtid; c < n_c; c += gthreads) {
        const size_t o = 9 * (size_t)c;
        double x[9], pp[9], w[9], wn[9];
        ld9(p.X + o, x); ld9(p.P + o, pp); ld9(p.W + o, w);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double s = 0.0;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) s += x[3 * i + jj] * H[3 * jj + j] + pp[3 * i + jj] * H[9 + 3 * jj + j];
                w[3 * i + j] -= s;
            }
#pragma unroll
        for (int i = 0; i < 9; ++i) wn[i] = 0.0;
        blk_times(w, T, 3, wn);
        st9(p.W + o, wn);
    }
    lob_stamp(p, 9);
}

inline int lobpcg_grid(int n_c) {
    int want = (n_c + LOB_THREADS - 1) / LOB_THREADS;
    const int cap = sm_count();   // cooperative launch: all CTAs co-resident (1 CTA / SM)
    if (want < 1) want = 1;
    return want < cap ? want : cap;
}

inline int launch_lobpcg_step(const LobpcgParams& p, cudaStream_t st) {
    LobpcgParams pp = p;
    void* args[] = {(void*)&pp};
    VB_CHECK(cudaLaunchCooperativeKernel((void*)lobpcg_step_kernel, dim3(lobpcg_grid(p.n_c)), dim3(LOB_THREADS), args, 0, st));
    return 0;
}

}  // namespace vb
