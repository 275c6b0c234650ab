// Translation stage (replaces vican/bipgo.py:434-481).
//
// The reference builds the 3E_raw x 3N incidence matrix J (rows k_t (x_t - x_c)) and solves
// with scipy's cg on J^T J or lsqr on J.  Here J is never formed:
//   * J^T J = (bipartite graph Laplacian with weights w_ct = sum k_t^2) (x) I_3;
//   * the iterations replay scipy's recurrences, stopping rules and (for cg) row-sum order
//     (cg.cuh, lsqr.cuh), because the reference result is a TRUNCATED
//     iterate, not the exact minimiser;
//   * all scalars (rho, alpha, beta, norms, stop flags) live on the device.
// This file: right-hand side J^T t~ and small helpers shared by the solvers.
#pragma once
#include <cub/cub.cuh>

#include "../../include/vican_b200.h"
#include "common.cuh"
#include "ingest.cuh"
#include "passes.cuh"
#include "rotation.cuh"

namespace vb {

constexpr int TR_THREADS = 256;
inline int tr_grid(int64_t n) { return (int)((n + TR_THREADS - 1) / TR_THREADS); }
inline int tr_warp_grid(int64_t n_warps) { return (int)((n_warps * 32 + TR_THREADS - 1) / TR_THREADS); }

// ------------------------------------------------------------------------------------ RHS
// thread per pair: g_p = sum k_t^2 d_e,  d_e = r_c^T t_cm + r_t^T q_m   (world rotation = r^T).
// The kernel is bound by the DEPTH of its dependent global loads (ncu: 46 warps per issue on the long scoreboard):
// pair_start -> raw_perm -> detection -> marker constant.  Two levels are removed where possible: the marker
// constants sit in shared memory (<= TR_MAX_MARKERS_SMEM markers), and detections that arrived sorted need no
// permutation (identity_perm: sorted position = raw index).
constexpr int TR_MAX_MARKERS_SMEM = 256;
__global__ void trans_pair_kernel(const int* __restrict__ raw_perm, const int* __restrict__ pair_start,
                                  const int* __restrict__ marker, const double* __restrict__ t_cm,
                                  const double* __restrict__ k_t, const double* __restrict__ marker_q_g,
                                  const double* __restrict__ r_c_pad, const double* __restrict__ r_t,
                                  const int* __restrict__ t_cam, const int* __restrict__ t_time, int64_t n_pairs,
                                  double* __restrict__ pair_g, double* __restrict__ d_sorted, int n_markers, int identity_perm) {
    __shared__ double sq[3 * TR_MAX_MARKERS_SMEM];
    const bool q_smem = n_markers <= TR_MAX_MARKERS_SMEM;
    if (q_smem)
        for (int i = threadIdx.x; i < 3 * n_markers; i += blockDim.x) sq[i] = marker_q_g[i];
    __syncthreads();
    const double* marker_q = q_smem ? sq : marker_q_g;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double Rc[9], Rt[9];
    const int64_t c = t_cam[p], t = t_time[p];
    // camera rotation: three 256-bit row loads from the padded copy (every lane reads another camera: one L1
    // wavefront per load and lane instead of nine); time rotation: neighbouring pairs share the node
    ld_row256(r_c_pad + GSTRIDE * c, Rc[0], Rc[1], Rc[2]);
    ld_row256(r_c_pad + GSTRIDE * c + 4, Rc[3], Rc[4], Rc[5]);
    ld_row256(r_c_pad + GSTRIDE * c + 8, Rc[6], Rc[7], Rc[8]);
#pragma unroll
    for (int i = 0; i < 9; ++i) Rt[i] = r_t[9 * t + i];
    double g0 = 0, g1 = 0, g2 = 0;
    for (int pos = pair_start[p]; pos < pair_start[p + 1]; ++pos) {
        const int64_t r = identity_perm ? pos : raw_perm[pos];
        const double tc[3] = {t_cm[3 * r], t_cm[3 * r + 1], t_cm[3 * r + 2]};
        const int64_t m = marker[r];
        const double q[3] = {marker_q[3 * m], marker_q[3 * m + 1], marker_q[3 * m + 2]};
        double a[3], b[3];
        mtv3(Rc, tc, a);
        mtv3(Rt, q, b);
        const double d0 = a[0] + b[0], d1 = a[1] + b[1], d2 = a[2] + b[2];
        const double w = k_t[r] * k_t[r];
        g0 += w * d0; g1 += w * d1; g2 += w * d2;
        if (d_sorted) { d_sorted[3 * (int64_t)pos] = d0; d_sorted[3 * (int64_t)pos + 1] = d1; d_sorted[3 * (int64_t)pos + 2] = d2; }
    }
    pair_g[3 * p] = g0; pair_g[3 * p + 1] = g1; pair_g[3 * p + 2] = g2;
}

// warp per node: out[v] = sign * sum_{i in segment} g[perm ? perm[i] : i]
__global__ void seg_sum3_kernel(const int* __restrict__ ptr, const int* __restrict__ perm, const double* __restrict__ g,
                                double sign, double* __restrict__ out, int64_t n_nodes) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_nodes) return;
    const int s = ptr[warp], e = ptr[warp + 1];
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = s + lane; i < e; i += 32) {
        const int64_t j = perm ? perm[i] : i;
        a0 += g[3 * j]; a1 += g[3 * j + 1]; a2 += g[3 * j + 2];
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { out[3 * warp] += sign * a0; out[3 * warp + 1] += sign * a1; out[3 * warp + 2] += sign * a2; }
}

// warp per camera: out[c] += sign * sum over the camera's (window, camera) runs of g[order[i]]
__global__ void cam_runs_sum3_kernel(const int* __restrict__ segptr, int64_t n_win, int64_t n_c, const int* __restrict__ order,
                                     const double* __restrict__ g, double sign, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_c) return;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int64_t wdw = 0; wdw < n_win; ++wdw) {
        const int s = segptr[wdw * n_c + c], e = segptr[wdw * n_c + c + 1];
        for (int i = s + lane; i < e; i += 32) {
            const int64_t j = order[i];
            a0 += g[3 * j]; a1 += g[3 * j + 1]; a2 += g[3 * j + 2];
        }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { out[3 * c] += sign * a0; out[3 * c + 1] += sign * a1; out[3 * c + 2] += sign * a2; }
}

// (the conjugate-gradient solver lives in cg.cuh; the helpers below serve the dense direct path)
template <int N>
__device__ __forceinline__ void block_atomic_sum(double (&v)[N], double* const (&dst)[N]) {
    __shared__ double sm[N][TR_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) sm[i][warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < TR_THREADS / 32; ++w) s += sm[threadIdx.x][w];
        atomicAdd(dst[threadIdx.x], s);
    }
}

// weighted degrees (diagonal of J^T J): dg_t = sum_row w, dg_c = sum_col w
__global__ void seg_sum1_kernel(const int* __restrict__ ptr, const int* __restrict__ perm, const double* __restrict__ w,
                                double* __restrict__ out, int64_t n_nodes) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_nodes) return;
    double a = 0;
    for (int i = ptr[warp] + lane; i < ptr[warp + 1]; i += 32) a += w[perm ? perm[i] : i];
    a = warp_sum(a);
    if (lane == 0) out[warp] = a;
}

}  // namespace vb
