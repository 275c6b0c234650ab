// Evaluation helpers of the reference as batched fp64 kernels: the step right
// after the solver in the notebook (main.ipynb cell 9) and the inner loop of the parity harness.
//   optimize_gauge_SO3 / optimize_gauge_SE3   vican/geometry.py:264-324
//   angle / distance_SO3                      vican/geometry.py:131-172
#pragma once
#include "common.cuh"

namespace vb {

constexpr int EV_THREADS = 256;
constexpr int EV_MAX_BLOCKS = 1024;

inline int ev_grid(int64_t n) {
    int64_t b = (n + EV_THREADS - 1) / EV_THREADS;
    return (int)(b < 1 ? 1 : (b > EV_MAX_BLOCKS ? EV_MAX_BLOCKS : b));
}

// partial[b][0..8]  = sum_i Ra_i^T Rb_i            (geometry.py:317)
// partial[b][9..11] = sum_i Rb_i^T (ta_i - tb_i)   (geometry.py:318)
// Fixed summation order (grid-stride per thread, shuffle tree, per-block slot): the result
// does not depend on scheduling.
__global__ void __launch_bounds__(EV_THREADS)
gauge_partial_kernel(const double* __restrict__ Ra, const double* __restrict__ ta, const double* __restrict__ Rb,
                     const double* __restrict__ tb, int64_t n, double* __restrict__ partial) {
    __shared__ double sm[EV_THREADS / 32][12];
    double acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double A[9], B[9], C[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { A[k] = Ra[9 * i + k]; B[k] = Rb[9 * i + k]; }
        mtm3(A, B, C);
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] += C[k];
        if (ta != nullptr) {
            const double d[3] = {ta[3 * i] - tb[3 * i], ta[3 * i + 1] - tb[3 * i + 1], ta[3 * i + 2] - tb[3 * i + 2]};
            double y[3];
            mtv3(B, d, y);
            acc[9] += y[0]; acc[10] += y[1]; acc[11] += y[2];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 12; ++k) sm[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int w = 0; w < EV_THREADS / 32; ++w) s += sm[w][threadIdx.x];
        partial[12 * (size_t)blockIdx.x + threadIdx.x] = s;
    }
}

// gauge_R = project_SO3(sum^T) (geometry.py:320-321: svd(sum.T), u diag(1,1,det(u vh)) vh);
// gauge_t = sum_t / n (geometry.py:322)
__global__ void gauge_finish_kernel(const double* __restrict__ partial, int n_blocks, int64_t n, double* __restrict__ gauge_R,
                                    double* __restrict__ gauge_t) {
    __shared__ double tot[12];
    if (threadIdx.x < 12) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; ++b) s += partial[12 * (size_t)b + threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double St[9], R[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) St[3 * i + j] = tot[3 * j + i];
        svd3_factors(St, R, nullptr, nullptr);   // U diag(1,1,det(U V^T)) V^T
#pragma unroll
        for (int k = 0; k < 9; ++k) gauge_R[k] = R[k];
        if (gauge_t != nullptr) {
            gauge_t[0] = tot[9] / (double)n; gauge_t[1] = tot[10] / (double)n; gauge_t[2] = tot[11] / (double)n;
        }
    }
}

// distance_SO3 (geometry.py:154-172) = angle(r1^T r2) in DEGREES with the reference's formula
// arccos(clip((trace - 1) / 2, -1, 1)) (geometry.py:150).  R2 == nullptr: angle(R1) (geometry.py:131-151).
__global__ void __launch_bounds__(EV_THREADS)
distance_so3_kernel(const double* __restrict__ R1, const double* __restrict__ R2, double* __restrict__ deg, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tr;
    if (R2 != nullptr) {
        tr = 0.0;
        // trace(R1^T R2) = sum_k R1[k] R2[k], accumulated diagonal entry by diagonal entry like the matrix product
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 3; ++r) s += R1[9 * i + 3 * r + j] * R2[9 * i + 3 * r + j];
            tr += s;
        }
    } else {
        tr = R1[9 * i] + R1[9 * i + 4] + R1[9 * i + 8];
    }
    double c = (tr - 1.0) / 2.0;
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    deg[i] = acos(c) * (180.0 / 3.14159265358979323846);
}

// (Rg, tg) @ (R_i, t_i) for ONE left transform and n poses  (cell 9: est = G.inv() @ pose_est[c])
__global__ void __launch_bounds__(EV_THREADS)
se3_left_compose_kernel(const double* __restrict__ Rg, const double* __restrict__ tg, const double* __restrict__ R,
                        const double* __restrict__ t, double* __restrict__ Ro, double* __restrict__ to, int64_t n, int round_f32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double A[9], B[9], C[9], a[3], b[3], c[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { A[k] = Rg[k]; B[k] = R[9 * i + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = tg[k]; b[k] = t[3 * i + k]; }
    if (round_f32) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { A[k] = (double)(float)A[k]; B[k] = (double)(float)B[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = (double)(float)a[k]; b[k] = (double)(float)b[k]; }
    }
    mm3(A, B, C);
    mv3(A, b, c);
#pragma unroll
    for (int k = 0; k < 9; ++k) Ro[9 * i + k] = round_f32 ? (double)(float)C[k] : C[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double v = c[k] + a[k]; to[3 * i + k] = round_f32 ? (double)(float)v : v; }
}

}  // namespace vb
