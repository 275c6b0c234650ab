// Edge passes of the matrix-free SE(3) Laplacian (the HBM-bound hot kernels).
//
// The reference forms the 3n_c x 3n_c power-graph matrix P Lambda_T P^T with an SpGEMM every
// outer iteration and multiplies by it with scipy's CSR kernels (vican/bipgo.py:273, :300,
// :318, :334).  Here nothing is formed.  One application of  P Lambda_T P^T  to a block
// vector X (n_c blocks of 3x3) is two streaming passes over the aggregated edge blocks:
//
//   time pass  (edges sorted by time node):  Z_t = sum_{e in t} B_e^T X_{c_e},  W_t = Lambda_T[t] Z_t
//   camera pass(edges sorted by camera):     Y_c = sum_{e in c} B_e   W_{t_e}
//
// Both are *gathers* on the far endpoint and segmented reductions on the owning endpoint, so
// no fp64 scatter-atomics per edge are needed (450 M atomics per pass at 50 M edges would be
// 3.5x slower than streaming the blocks; shared-memory fp64 atomics are CAS loops).
//
// Lane mapping: a warp owns one segment (a time node, or a tile of one camera's edges) and
// walks it three edges per round with 27 active lanes = 3 edges x 9 block entries.  Lane
// (q, r) loads entry r of edge q's block -> the [E][9] block array is read as one flat,
// fully coalesced stream.  The same lane loads entry r of the gathered 3x3 node block (one
// 72-byte record per edge, covered by 9 adjacent lanes of ONE load instruction, ~1.3 L1
// wavefronts per edge instead of 9 with a lane-per-edge mapping); the row it needs is
// fetched from its 3 neighbour lanes with shuffles.
//
// Algorithmic bytes per edge: 72 (block) + 4 (index) = 76 B  (SURVEY.md 8d).
#pragma once
#include "common.cuh"

namespace vb {

constexpr int PASS_THREADS = 256;
constexpr int CHUNK = 24;   // edge indices fetched per chunk (one per lane)
constexpr int UNR = 4;      // rounds (of 3 edges) whose loads are issued back to back

// acc[j] partial sums over edges [s, e) for lane (q, r):
//   TR  (time pass): acc[j] += B_e[k][i] * G_e[k][j]   with r = 3k + i   -> (B^T G)[i][j]
//   !TR (cam pass) : acc[j] += B_e[i][k] * G_e[k][j]   with r = 3i + k   -> (B   G)[i][j]
template <bool TR>
__device__ __forceinline__ void edge_accumulate(const double* __restrict__ B, const int* __restrict__ idx,
                                                const double* __restrict__ G, int s, int e, int lane,
                                                uint64_t pol_stream, uint64_t pol_keep, double& a0, double& a1,
                                                double& a2) {
    const int q = lane / 9;
    const int r = lane - 9 * q;
    const int krow = TR ? (r / 3) : (r % 3);
    const int src = 9 * q + 3 * krow;
    for (int cb = s; cb < e; cb += CHUNK) {
        int my_idx = 0;
        if (lane < CHUNK && cb + lane < e) my_idx = ld_stream(idx + cb + lane, pol_stream);
#pragma unroll
        for (int g = 0; g < CHUNK / (3 * UNR); ++g) {
            if (cb + 3 * UNR * g >= e) break;   // warp-uniform
            double b[UNR], x[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int off = 3 * UNR * g + 3 * u + q;
                const int ed = cb + off;
                const bool on = (q < 3) && (ed < e);
                const int node = __shfl_sync(FULL, my_idx, off & 31);
                b[u] = on ? ld_stream(B + 9 * (size_t)ed + r, pol_stream) : 0.0;
                x[u] = on ? ld_keep(G + 9 * (size_t)node + r, pol_keep) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const double x0 = shfl(x[u], src), x1 = shfl(x[u], src + 1), x2 = shfl(x[u], src + 2);
                a0 = fma(b[u], x0, a0);
                a1 = fma(b[u], x1, a1);
                a2 = fma(b[u], x2, a2);
            }
        }
    }
}

// Sum the 9 partial rows that belong to the same output row.  On return the totals live in
//   TR : lanes 0,1,2   (lane = output row)       !TR : lanes 0,3,6 (lane/3 = output row)
template <bool TR>
__device__ __forceinline__ void edge_reduce(double& a0, double& a1, double& a2) {
    a0 += shfl_down(a0, 9) + shfl_down(a0, 18);
    a1 += shfl_down(a1, 9) + shfl_down(a1, 18);
    a2 += shfl_down(a2, 9) + shfl_down(a2, 18);
    constexpr int d1 = TR ? 3 : 1, d2 = TR ? 6 : 2;
    a0 += shfl_down(a0, d1) + shfl_down(a0, d2);
    a1 += shfl_down(a1, d1) + shfl_down(a1, d2);
    a2 += shfl_down(a2, d1) + shfl_down(a2, d2);
}

// MODE 0: out_t = Lambda_T[t] * Z_t  (L-apply / primal multiply, bipgo.py:300 first half)
// MODE 1: out_t = Z_t               (dual gather Y = P^T r_c, bipgo.py:318)
template <int MODE>
__global__ void __launch_bounds__(PASS_THREADS)
pass_time_kernel(const int* __restrict__ rowptr, const int* __restrict__ cam, const double* __restrict__ B,
                 const double* __restrict__ X, const double* __restrict__ lamT, double* __restrict__ out, int n_t) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    for (int node = warp; node < n_t; node += nwarps) {
        const int s = __ldg(rowptr + node), e = __ldg(rowptr + node + 1);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        edge_accumulate<true>(B, cam, X, s, e, lane, pf, pl, a0, a1, a2);
        edge_reduce<true>(a0, a1, a2);
        if (MODE == 0) {
            // every lane gets the full Z (rows live in lanes 0..2)
            double z[9];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                z[3 * i] = shfl(a0, i);
                z[3 * i + 1] = shfl(a1, i);
                z[3 * i + 2] = shfl(a2, i);
            }
            if (lane < 3) {
                const double* L = lamT + 9 * (size_t)node + 3 * lane;
                const double l0 = L[0], l1 = L[1], l2 = L[2];
                double* o = out + 9 * (size_t)node + 3 * lane;
                o[0] = l0 * z[0] + l1 * z[3] + l2 * z[6];
                o[1] = l0 * z[1] + l1 * z[4] + l2 * z[7];
                o[2] = l0 * z[2] + l1 * z[5] + l2 * z[8];
            }
        } else {
            if (lane < 3) {
                double* o = out + 9 * (size_t)node + 3 * lane;
                o[0] = a0; o[1] = a1; o[2] = a2;
            }
        }
    }
}

// Y_c += sum over one tile (a run of edges of a single camera) of B_e W_{t_e}.
// Tiles of one camera are combined with 9 fp64 atomics per tile (not per edge).
__global__ void __launch_bounds__(PASS_THREADS)
pass_cam_kernel(const int* __restrict__ tile_cam, const int* __restrict__ tile_start,
                const int* __restrict__ tile_end, const int* __restrict__ tidx, const double* __restrict__ B,
                const double* __restrict__ W, double* __restrict__ Y, int n_tiles) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    for (int tile = warp; tile < n_tiles; tile += nwarps) {
        const int c = __ldg(tile_cam + tile), s = __ldg(tile_start + tile), e = __ldg(tile_end + tile);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        edge_accumulate<false>(B, tidx, W, s, e, lane, pf, pl, a0, a1, a2);
        edge_reduce<false>(a0, a1, a2);
        if (lane == 0 || lane == 3 || lane == 6) {
            double* y = Y + 9 * (size_t)c + lane;   // row lane/3 starts at 3*(lane/3) = lane
            atomicAdd(y, a0);
            atomicAdd(y + 1, a1);
            atomicAdd(y + 2, a2);
        }
    }
}

inline int pass_grid(int64_t n_segments) {
    const int64_t warps_per_block = PASS_THREADS / 32;
    const int64_t want = (n_segments + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)sm_count() * 8;   // <= 8 resident CTAs of 256 threads per SM
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

inline int launch_pass_time(int mode, const int* rowptr, const int* cam, const double* B, const double* X,
                            const double* lamT, double* out, int64_t n_t, cudaStream_t st) {
    if (n_t <= 0) return 0;
    const int grid = pass_grid(n_t);
    if (mode == 0)
        pass_time_kernel<0><<<grid, PASS_THREADS, 0, st>>>(rowptr, cam, B, X, lamT, out, (int)n_t);
    else
        pass_time_kernel<1><<<grid, PASS_THREADS, 0, st>>>(rowptr, cam, B, X, lamT, out, (int)n_t);
    VB_KERNEL_CHECK();
    return 0;
}

inline int launch_pass_cam(const int* tile_cam, const int* tile_start, const int* tile_end, const int* tidx,
                           const double* B, const double* W, double* Y, int64_t n_tiles, cudaStream_t st) {
    if (n_tiles <= 0) return 0;
    pass_cam_kernel<<<pass_grid(n_tiles), PASS_THREADS, 0, st>>>(tile_cam, tile_start, tile_end, tidx, B, W, Y,
                                                                 (int)n_tiles);
    VB_KERNEL_CHECK();
    return 0;
}

}  // namespace vb
