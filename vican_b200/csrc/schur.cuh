// Dense "direct" translation solve for small camera sets (north_star (3): cuSOLVER-free dense
// block path for the 'direct' option on small graphs).
//
// The reference's normal equations are J^T J x = J^T t~ (vican/bipgo.py:471-478) with
// J^T J = L_w (x) I_3, L_w = [D_c  -W; -W^T  D_t] the weighted bipartite graph Laplacian
// (w_ct = sum k_t^2 over the detections of the pair).  D_t is diagonal, so the time nodes are
// eliminated in closed form and only the n_c x n_c camera Schur complement
//     S = D_c - W D_t^-1 W^T,      g = b_c + W D_t^-1 b_t      (3 right-hand sides)
// is dense.  L_w has the constant null vector (global translation): the last camera is grounded
// (x = 0), the leading (n_c-1)^2 block is factored with a blocked right-looking Cholesky written
// here (32 x 32 tiles in shared memory, fp64 FMA), time nodes are back-substituted, and the mean
// over all N nodes is removed, which gives the minimum-norm minimiser -- the point scipy's
// cg / lsqr iterate towards from x0 = 0 (they stop ~1e-5 short of it).
#pragma once
#include "../../include/vican_b200.h"
#include "common.cuh"
#include "ingest.cuh"
#include "rotation.cuh"
#include "translation.cuh"

namespace vb {

constexpr int CH_NB = 32;          // Cholesky tile
constexpr int SCHUR_MAX_NC = 8192;

struct SchurWork {
    double *S, *G, *dg_c, *dg_t, *sums;
    int* flag;
    int64_t bytes;
};

inline SchurWork carve_schur(void* base, int64_t n_c, int64_t n_t) {
    SchurWork w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t nd) {
        double* r = (double*)(p + off);
        off += align256(nd * (int64_t)sizeof(double));
        return r;
    };
    w.S = take(n_c * n_c); w.G = take(3 * n_c); w.dg_c = take(n_c); w.dg_t = take(n_t); w.sums = take(8);
    w.flag = (int*)take(1);
    w.bytes = off;
    return w;
}

// S = diag(dg_c) (dense, zeroed by the caller), G = b_c
__global__ void schur_init_kernel(const double* __restrict__ dg_c, const double* __restrict__ b_c, double* __restrict__ S,
                                  double* __restrict__ G, int64_t n_c) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    S[c * n_c + c] = dg_c[c];
    G[3 * c] = b_c[3 * c]; G[3 * c + 1] = b_c[3 * c + 1]; G[3 * c + 2] = b_c[3 * c + 2];
}

// warp per time node: S[c_i][c_j] -= w_i w_j / d_t for all pairs of its edges, G[c_i] += w_i b_t / d_t
__global__ void schur_build_kernel(const int* __restrict__ rowptr, const int* __restrict__ cam, const double* __restrict__ w,
                                   const double* __restrict__ dg_t, const double* __restrict__ b_t, double* __restrict__ S,
                                   double* __restrict__ G, int64_t n_c, int64_t n_t) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_t) return;
    const int s = rowptr[t], e = rowptr[t + 1];
    const double inv = 1.0 / dg_t[t];
    const double b0 = b_t[3 * t] * inv, b1 = b_t[3 * t + 1] * inv, b2 = b_t[3 * t + 2] * inv;
    for (int i = s + lane; i < e; i += 32) {
        const int64_t ci = cam[i];
        const double wi = w[i];
        atomicAdd(G + 3 * ci, wi * b0); atomicAdd(G + 3 * ci + 1, wi * b1); atomicAdd(G + 3 * ci + 2, wi * b2);
    }
    const int64_t deg = e - s;
    for (int64_t pq = lane; pq < deg * deg; pq += 32) {
        const int i = s + (int)(pq / deg), j = s + (int)(pq % deg);
        atomicAdd(S + (int64_t)cam[i] * n_c + cam[j], -(w[i] * w[j]) * inv);
    }
}

// ---- blocked right-looking Cholesky of the leading m x m block of S (row-major, ld), lower part
__global__ void __launch_bounds__(CH_NB* CH_NB) chol_diag_kernel(double* __restrict__ S, int64_t ld, int64_t k0, int nb, const double* __restrict__ diag0, int* flag) {
    __shared__ double A[CH_NB][CH_NB + 1];
    const int i = threadIdx.y, j = threadIdx.x;
    A[i][j] = (i < nb && j < nb) ? S[(k0 + i) * ld + k0 + j] : (i == j ? 1.0 : 0.0);
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        if (i == k && j == k) {
            // pivot relative to the original diagonal entry (weighted degree): an exactly singular
            // Schur complement (disconnected graph) leaves pivots of rounding size, not zeros
            const double d = A[k][k];
            if (!(d > 1e-12 * diag0[k0 + k])) { *flag = 1; A[k][k] = 1.0; } else A[k][k] = sqrt(d);
        }
        __syncthreads();
        if (j == k && i > k) A[i][k] /= A[k][k];
        __syncthreads();
        if (i > k && j > k && j <= i) A[i][j] -= A[i][k] * A[j][k];
        __syncthreads();
    }
    if (i < nb && j < nb) S[(k0 + i) * ld + k0 + j] = (j <= i) ? A[i][j] : 0.0;
}

// rows below the diagonal block: L_ik = A_ik L_kk^-T  (thread per row, forward substitution)
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ S, int64_t ld, int64_t k0, int nb, int64_t m) {
    __shared__ double Lk[CH_NB][CH_NB + 1];
    for (int idx = threadIdx.x; idx < CH_NB * CH_NB; idx += blockDim.x) {
        const int i = idx / CH_NB, j = idx % CH_NB;
        Lk[i][j] = (i < nb && j < nb) ? S[(k0 + i) * ld + k0 + j] : 0.0;
    }
    __syncthreads();
    const int64_t r = k0 + nb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double x[CH_NB];
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) x[j] = (j < nb) ? S[r * ld + k0 + j] : 0.0;
#pragma unroll
    for (int j = 0; j < CH_NB; ++j) {
        if (j < nb) {
            double v = x[j];
#pragma unroll
            for (int q = 0; q < CH_NB; ++q)
                if (q < j) v -= x[q] * Lk[j][q];
            x[j] = v / Lk[j][j];
        }
    }
#pragma unroll
    for (int j = 0; j < CH_NB; ++j)
        if (j < nb) S[r * ld + k0 + j] = x[j];
}

// trailing update A_ij -= L_ik L_jk^T over 32 x 32 tiles with i >= j
__global__ void __launch_bounds__(CH_NB* CH_NB) chol_update_kernel(double* __restrict__ S, int64_t ld, int64_t k0, int nb, int64_t m) {
    if (blockIdx.x > blockIdx.y) return;   // lower triangle of tiles only
    __shared__ double Li[CH_NB][CH_NB + 1], Lj[CH_NB][CH_NB + 1];
    const int64_t base = k0 + nb;
    const int64_t ri = base + (int64_t)blockIdx.y * CH_NB + threadIdx.y;
    const int64_t rj = base + (int64_t)blockIdx.x * CH_NB + threadIdx.y;
    Li[threadIdx.y][threadIdx.x] = (ri < m && threadIdx.x < nb) ? S[ri * ld + k0 + threadIdx.x] : 0.0;
    Lj[threadIdx.y][threadIdx.x] = (rj < m && threadIdx.x < nb) ? S[rj * ld + k0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int64_t cj = base + (int64_t)blockIdx.x * CH_NB + threadIdx.x;
    if (ri >= m || cj >= m || cj > ri) return;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < CH_NB; ++k) acc = fma(Li[threadIdx.y][k], Lj[threadIdx.x][k], acc);
    S[ri * ld + cj] -= acc;
}

// L y = g, L^T x = y for 3 right-hand sides, one CTA (the small-graph path: m <= a few thousand)
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double* __restrict__ S, int64_t ld, int64_t m, double* __restrict__ G) {
    __shared__ double piv[3];
    const int tid = threadIdx.x, nth = blockDim.x;
    for (int64_t j = 0; j < m; ++j) {                 // forward: column-oriented
        if (tid < 3) { const double v = G[3 * j + tid] / S[j * ld + j]; G[3 * j + tid] = v; piv[tid] = v; }
        __syncthreads();
        for (int64_t i = j + 1 + tid; i < m; i += nth) {
            const double l = S[i * ld + j];
            G[3 * i] -= l * piv[0]; G[3 * i + 1] -= l * piv[1]; G[3 * i + 2] -= l * piv[2];
        }
        __syncthreads();
    }
    for (int64_t j = m - 1; j >= 0; --j) {            // backward with L^T: row j of L holds column j of L^T
        if (tid < 3) { const double v = G[3 * j + tid] / S[j * ld + j]; G[3 * j + tid] = v; piv[tid] = v; }
        __syncthreads();
        for (int64_t i = tid; i < j; i += nth) {
            const double l = S[j * ld + i];
            G[3 * i] -= l * piv[0]; G[3 * i + 1] -= l * piv[1]; G[3 * i + 2] -= l * piv[2];
        }
        __syncthreads();
    }
}

// x_c = G (grounded camera = 0); x_t = (b_t + sum_e w_e x_c[c_e]) / d_t; accumulates the sums for the mean
__global__ void schur_copy_xc_kernel(const double* __restrict__ G, double* __restrict__ x_c, int64_t n_c, int64_t m, double* sums) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    if (c < n_c) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { v[k] = (c < m) ? G[3 * c + k] : 0.0; x_c[3 * c + k] = v[k]; }
    }
    double* const dst[3] = {sums, sums + 1, sums + 2};
    block_atomic_sum<3>(v, dst);
}

__global__ void schur_back_kernel(const int* __restrict__ rowptr, const int* __restrict__ cam, const double* __restrict__ w,
                                  const double* __restrict__ dg_t, const double* __restrict__ b_t, const double* __restrict__ x_c,
                                  double* __restrict__ x_t, int64_t n_t, double* sums) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double v[3] = {0.0, 0.0, 0.0};
    if (t < n_t) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int i = rowptr[t] + lane; i < rowptr[t + 1]; i += 32) {
            const int64_t c = cam[i];
            const double ww = w[i];
            a0 += ww * x_c[3 * c]; a1 += ww * x_c[3 * c + 1]; a2 += ww * x_c[3 * c + 2];
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        if (lane == 0) {
            const double inv = 1.0 / dg_t[t];
            v[0] = (b_t[3 * t] + a0) * inv; v[1] = (b_t[3 * t + 1] + a1) * inv; v[2] = (b_t[3 * t + 2] + a2) * inv;
            x_t[3 * t] = v[0]; x_t[3 * t + 1] = v[1]; x_t[3 * t + 2] = v[2];
        }
    }
    double* const dst[3] = {sums + 3, sums + 4, sums + 5};
    block_atomic_sum<3>(v, dst);
}

__global__ void schur_center_kernel(double* __restrict__ x, int64_t n, const double* __restrict__ sums, double inv_N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int k = (int)(i % 3);
    x[i] -= (sums[k] + sums[3 + k]) * inv_N;
}

inline int trans_schur_direct(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t,
                              void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    if (n_c < 2 || n_c > SCHUR_MAX_NC || n_t < 1) return VB_STATUS_BAD_ARGUMENT;
    SchurWork w = carve_schur(workspace, n_c, n_t);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    VB_CHECK(cudaMemsetAsync(w.S, 0, n_c * n_c * sizeof(double), st));
    VB_CHECK(cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), st));
    VB_CHECK(cudaMemsetAsync(w.flag, 0, sizeof(int), st));
    seg_sum1_kernel<<<tr_warp_grid(n_t), TR_THREADS, 0, st>>>(g->t_rowptr, nullptr, g->t_w, w.dg_t, n_t);
    cam_runs_sum_kernel<<<tr_warp_grid(n_c), TR_THREADS, 0, st>>>(g->c_segptr, g->n_windows, n_c, nullptr, g->c_w, w.dg_c);
    schur_init_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(w.dg_c, rhs_c, w.S, w.G, n_c);
    schur_build_kernel<<<tr_warp_grid(n_t), TR_THREADS, 0, st>>>(g->t_rowptr, g->t_cam, g->t_w, w.dg_t, rhs_t, w.S, w.G, n_c, n_t);
    VB_KERNEL_CHECK();
    const int64_t m = n_c - 1;   // grounded: last camera
    for (int64_t k0 = 0; k0 < m; k0 += CH_NB) {
        const int nb = (int)((m - k0) < CH_NB ? (m - k0) : CH_NB);
        chol_diag_kernel<<<1, dim3(CH_NB, CH_NB), 0, st>>>(w.S, n_c, k0, nb, w.dg_c, w.flag);
        const int64_t rest = m - k0 - nb;
        if (rest > 0) {
            chol_panel_kernel<<<(int)((rest + 127) / 128), 128, 0, st>>>(w.S, n_c, k0, nb, m);
            const int tiles = (int)((rest + CH_NB - 1) / CH_NB);
            chol_update_kernel<<<dim3(tiles, tiles), dim3(CH_NB, CH_NB), 0, st>>>(w.S, n_c, k0, nb, m);
        }
    }
    VB_KERNEL_CHECK();
    chol_solve_kernel<<<1, 1024, 0, st>>>(w.S, n_c, m, w.G);
    schur_copy_xc_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(w.G, x_c, n_c, m, w.sums);
    schur_back_kernel<<<tr_warp_grid(n_t), TR_THREADS, 0, st>>>(g->t_rowptr, g->t_cam, g->t_w, w.dg_t, rhs_t, x_c, x_t, n_t, w.sums);
    const double inv_N = 1.0 / (double)(n_c + n_t);
    schur_center_kernel<<<tr_grid(3 * n_c), TR_THREADS, 0, st>>>(x_c, n_c, w.sums, inv_N);
    schur_center_kernel<<<tr_grid(3 * n_t), TR_THREADS, 0, st>>>(x_t, n_t, w.sums, inv_N);
    VB_KERNEL_CHECK();
    int h_flag = 0;
    VB_CHECK(cudaMemcpyAsync(&h_flag, w.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_CHECK(cudaStreamSynchronize(st));
    return h_flag ? VB_STATUS_SINGULAR : VB_STATUS_OK;
}

}  // namespace vb
