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
#include <stdlib.h>
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

inline int launch_pass_time_v1(int mode, const int* rowptr, const int* cam, const double* B, const double* X,
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

inline int launch_pass_cam_v1(const int* tile_cam, const int* tile_start, const int* tile_end, const int* tidx,
                           const double* B, const double* W, double* Y, int64_t n_tiles, cudaStream_t st) {
    if (n_tiles <= 0) return 0;
    pass_cam_kernel<<<pass_grid(n_tiles), PASS_THREADS, 0, st>>>(tile_cam, tile_start, tile_end, tidx, B, W, Y,
                                                                 (int)n_tiles);
    VB_KERNEL_CHECK();
    return 0;
}


// =====================================================================================
// v2: per-warp TMA staging.  Profiling v1 (profiles/r1_passes_v1.md) showed both passes
// latency-bound: every warp walked a chain of ~10 dependent DRAM round trips per node
// (row pointer -> index chunk -> blocks/gather, per 24-edge chunk) with DRAM at 30-46 %.
// Here each warp owns a double-buffered shared-memory stage and an mbarrier pair; one elected
// lane streams the NEXT work item (<= 64 edges: their blocks and indices, two
// cp.async.bulk copies, evict-first) while the warp consumes the current one from shared
// memory.  Only the gathered node blocks still travel through LDG (L2 / L1 hits), 8 in flight
// per lane.  No CTA-level synchronisation at all: warps are independent pipelines.
// =====================================================================================
constexpr int PASS_WARPS = PASS_THREADS / 32;
constexpr int ITEM_EDGES = 64;
constexpr int BUF_B_BYTES = (ITEM_EDGES + 2) * 72;                    // 4752, multiple of 16
constexpr int BUF_I_BYTES = (ITEM_EDGES + 8) * 4;                     // 288
constexpr int BUF_BYTES = ((BUF_B_BYTES + BUF_I_BYTES + 127) / 128) * 128;   // 5120
constexpr int PASS_SMEM = PASS_WARPS * 2 * BUF_BYTES + PASS_WARPS * 2 * 8;
constexpr int GUNR = 8;                                               // rounds per gather group

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// One group of NR rounds (3 edges each) of the current item.  Lane (q, r) multiplies its block
// entry with one row of the gathered 3x3 block.  It loads ONE entry of that row itself and
// takes the other two from its two neighbour lanes, accumulating "rotated" partial sums
// (acc_d belongs to output column (rot + d) mod 3) -- 2 instead of 3 fp64 shuffles per round;
// the LSU/shuffle return path is what bounds this kernel (profiles/r1_passes_v2.md).
template <int NR>
__device__ __forceinline__ void edge_group(const double* __restrict__ sB, const int* __restrict__ sI,
                                           const double* __restrict__ G, int offB, int offI, int g0, int n_e,
                                           int q, int r, int gsel, int src1, int src2, uint64_t pl, double& c0,
                                           double& c1, double& c2) {
    double bv[NR], xv[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        const int off = g0 + 3 * u + q;
        const bool on = (q < 3) && (off < n_e);
        bv[u] = 0.0; xv[u] = 0.0;
        if (on) {
            const int node = sI[offI + off];
            bv[u] = sB[9 * (offB + off) + r];
            xv[u] = ld_keep(G + 9 * (size_t)node + gsel, pl);
        }
    }
#pragma unroll
    for (int u = 0; u < NR; ++u) {
        const double x1 = shfl(xv[u], src1), x2 = shfl(xv[u], src2);
        c0 = fma(bv[u], xv[u], c0);
        c1 = fma(bv[u], x1, c1);
        c2 = fma(bv[u], x2, c2);
    }
}

// MODE 0: out_t = Lambda_T[t] * sum B^T X   MODE 1: out_t = sum B^T X   MODE 2: Y_c += sum B W (atomics per tile)
template <int MODE>
__global__ void __launch_bounds__(PASS_THREADS, 2)
edge_pass_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node, const int* __restrict__ idx,
                 const double* __restrict__ B, const double* __restrict__ G, const double* __restrict__ lamT,
                 double* __restrict__ out, int n_seg) {
    extern __shared__ __align__(128) unsigned char pass_smem[];
    constexpr bool TR = (MODE != 2);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * PASS_WARPS + wib;
    const int nwarps = gridDim.x * PASS_WARPS;
    unsigned char* wbuf = pass_smem + (size_t)wib * 2 * BUF_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(pass_smem + (size_t)PASS_WARPS * 2 * BUF_BYTES) + 2 * wib;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (warp >= n_seg) return;

    const uint64_t pf = policy_evict_first(), pl = policy_evict_last();
    const int q = lane / 9, r = lane - 9 * q;
    // TR : r = 3k + i (block entry B[k][i]),  gathers X[k][i],   row mates = lanes 9q + 3k + *
    // !TR: r = 3i + k (block entry B[i][k]),  gathers W[k][i],   row mates = lanes 9q + 3* + k
    const int hi = r / 3, lo = r - 3 * hi;
    const int rot = TR ? lo : hi;                       // = output row of this lane
    const int gsel = TR ? r : (3 * lo + hi);
    const int m1 = (rot + 1) % 3, m2 = (rot + 2) % 3;
    const int src1 = TR ? (9 * q + 3 * hi + m1) : (9 * q + 3 * m1 + lo);
    const int src2 = TR ? (9 * q + 3 * hi + m2) : (9 * q + 3 * m2 + lo);

    auto issue = [&](int a, int b, int buf) {   // elected lane only
        const int a0 = a & ~1, b1 = (b + 1) & ~1;
        const int a0i = a & ~3, b1i = (b + 3) & ~3;
        const uint32_t nbB = (uint32_t)(b1 - a0) * 72u, nbI = (uint32_t)(b1i - a0i) * 4u;
        unsigned char* dst = wbuf + buf * BUF_BYTES;
        mbar_expect_tx(&bars[buf], nbB + nbI);
        bulk_g2s(dst, B + 9 * (size_t)a0, nbB, &bars[buf], pf);
        bulk_g2s(dst + BUF_B_BYTES, idx + a0i, nbI, &bars[buf], pf);
    };

    int cseg = warp;
    int cs = __ldg(seg_ptr + cseg), ce = __ldg(seg_ptr + cseg + 1);
    int ca = cs, cb = min(cs + ITEM_EDGES, ce);
    int nseg = cseg + nwarps, ns = 0, ne = 0;
    if (nseg < n_seg) { ns = __ldg(seg_ptr + nseg); ne = __ldg(seg_ptr + nseg + 1); }
    uint32_t phase0 = 0, phase1 = 0;
    int n_issued = 0, cur_buf = 0;
    bool cur_issued = false;
    if (cb > ca) {
        if (lane == 0) issue(ca, cb, 0);
        cur_issued = true; cur_buf = 0; n_issued = 1;
    }
    // Lambda_T row for the epilogue of the current segment, fetched early (latency hidden)
    double lam0 = 0.0, lam1 = 0.0, lam2 = 0.0;
    if (MODE == 0 && lane < 3) {
        const double* L = lamT + 9 * (size_t)cseg + 3 * lane;
        lam0 = L[0]; lam1 = L[1]; lam2 = L[2];
    }
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (;;) {
        // ---- look ahead: next chunk of this segment, or the head of the next segment
        bool has_next = true, next_new_seg = false;
        int na = 0, nb = 0;
        if (cb < ce) { na = cb; nb = min(cb + ITEM_EDGES, ce); }
        else if (nseg < n_seg) { na = ns; nb = min(ns + ITEM_EDGES, ne); next_new_seg = true; }
        else has_next = false;
        bool next_issued = false;
        int next_buf = 0;
        if (has_next && nb > na) {
            next_buf = n_issued & 1;
            __syncwarp();   // every lane is done reading that buffer (item before the current one)
            if (lane == 0) issue(na, nb, next_buf);
            next_issued = true;
            ++n_issued;
        }
        // ---- consume the current item from shared memory
        if (cur_issued) {
            if (cur_buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; }
            else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
            const unsigned char* bufp = wbuf + cur_buf * BUF_BYTES;
            const double* sB = reinterpret_cast<const double*>(bufp);
            const int* sI = reinterpret_cast<const int*>(bufp + BUF_B_BYTES);
            const int offB = ca - (ca & ~1), offI = ca - (ca & ~3);
            const int n_e = cb - ca;
            int g0 = 0;
            for (; g0 + 3 * 8 <= n_e; g0 += 3 * 8) edge_group<8>(sB, sI, G, offB, offI, g0, n_e, q, r, gsel, src1, src2, pl, c0, c1, c2);
            if (g0 + 3 * 4 <= n_e) { edge_group<4>(sB, sI, G, offB, offI, g0, n_e, q, r, gsel, src1, src2, pl, c0, c1, c2); g0 += 12; }
            if (g0 + 3 * 2 <= n_e) { edge_group<2>(sB, sI, G, offB, offI, g0, n_e, q, r, gsel, src1, src2, pl, c0, c1, c2); g0 += 6; }
            if (g0 + 3 <= n_e) { edge_group<1>(sB, sI, G, offB, offI, g0, n_e, q, r, gsel, src1, src2, pl, c0, c1, c2); g0 += 3; }
            if (g0 < n_e) edge_group<1>(sB, sI, G, offB, offI, g0, n_e, q, r, gsel, src1, src2, pl, c0, c1, c2);
        }
        // ---- segment finished: reduce (rotated sums share their rotation within a row), un-rotate, emit
        if (cb >= ce) {
            edge_reduce<TR>(c0, c1, c2);
            // final lanes: TR -> 0,1,2 (rot = lane) ; !TR -> 0,3,6 (rot = lane / 3)
            const double a0 = (rot == 0) ? c0 : ((rot == 1) ? c2 : c1);
            const double a1 = (rot == 0) ? c1 : ((rot == 1) ? c0 : c2);
            const double a2 = (rot == 0) ? c2 : ((rot == 1) ? c1 : c0);
            if (MODE == 0) {
                double z[9];
#pragma unroll
                for (int i = 0; i < 3; ++i) { z[3 * i] = shfl(a0, i); z[3 * i + 1] = shfl(a1, i); z[3 * i + 2] = shfl(a2, i); }
                if (lane < 3) {
                    double* o = out + 9 * (size_t)cseg + 3 * lane;
                    o[0] = lam0 * z[0] + lam1 * z[3] + lam2 * z[6];
                    o[1] = lam0 * z[1] + lam1 * z[4] + lam2 * z[7];
                    o[2] = lam0 * z[2] + lam1 * z[5] + lam2 * z[8];
                }
            } else if (MODE == 1) {
                if (lane < 3) { double* o = out + 9 * (size_t)cseg + 3 * lane; o[0] = a0; o[1] = a1; o[2] = a2; }
            } else {
                if (lane == 0 || lane == 3 || lane == 6) {
                    double* y = out + 9 * (size_t)__ldg(seg_node + cseg) + lane;
                    atomicAdd(y, a0); atomicAdd(y + 1, a1); atomicAdd(y + 2, a2);
                }
            }
            c0 = 0.0; c1 = 0.0; c2 = 0.0;
        }
        if (!has_next) break;
        // ---- advance
        if (next_new_seg) {
            cseg = nseg; cs = ns; ce = ne;
            nseg += nwarps;
            if (nseg < n_seg) { ns = __ldg(seg_ptr + nseg); ne = __ldg(seg_ptr + nseg + 1); }
            if (MODE == 0 && lane < 3) {
                const double* L = lamT + 9 * (size_t)cseg + 3 * lane;
                lam0 = L[0]; lam1 = L[1]; lam2 = L[2];
            }
        }
        ca = na; cb = nb; cur_issued = next_issued; cur_buf = next_buf;
    }
}

inline int pass_grid_v2(int64_t n_segments) {
    const int64_t want = (n_segments + PASS_WARPS - 1) / PASS_WARPS;
    const int64_t cap = (int64_t)sm_count() * 2;   // 2 resident CTAs per SM (82 KB smem each): persistent, one wave
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int MODE>
inline int launch_edge_pass(const int* seg_ptr, const int* seg_node, const int* idx, const double* B, const double* G,
                            const double* lamT, double* out, int64_t n_seg, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        VB_CHECK(cudaFuncSetAttribute(edge_pass_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PASS_SMEM));
        attr_set = true;
    }
    edge_pass_kernel<MODE><<<pass_grid_v2(n_seg), PASS_THREADS, PASS_SMEM, st>>>(seg_ptr, seg_node, idx, B, G, lamT, out, (int)n_seg);
    VB_KERNEL_CHECK();
    return 0;
}

inline int pass_impl() {
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("VB_PASS_IMPL");
        impl = (e && e[0] == '1') ? 1 : 2;
    }
    return impl;
}

// NOTE (v2): idx must be readable up to index ((E+3)&~3)-1 and B up to edge ((E+1)&~1)-1
// (the bulk copies are 16-byte granular); the ingestion allocates that padding.
inline int launch_pass_time(int mode, const int* rowptr, const int* cam, const double* B, const double* X,
                            const double* lamT, double* out, int64_t n_t, cudaStream_t st) {
    if (n_t <= 0) return 0;
    if (pass_impl() == 1) return launch_pass_time_v1(mode, rowptr, cam, B, X, lamT, out, n_t, st);
    if (mode == 0) return launch_edge_pass<0>(rowptr, nullptr, cam, B, X, lamT, out, n_t, st);
    return launch_edge_pass<1>(rowptr, nullptr, cam, B, X, lamT, out, n_t, st);
}

// tile_start must carry a sentinel: tile_start[n_tiles] = E (tiles are contiguous)
inline int launch_pass_cam(const int* tile_cam, const int* tile_start, const int* tile_end, const int* tidx,
                           const double* B, const double* W, double* Y, int64_t n_tiles, cudaStream_t st) {
    if (n_tiles <= 0) return 0;
    if (pass_impl() == 1) return launch_pass_cam_v1(tile_cam, tile_start, tile_end, tidx, B, W, Y, n_tiles, st);
    return launch_edge_pass<2>(tile_start, tile_cam, tidx, B, W, nullptr, Y, n_tiles, st);
}

}  // namespace vb
