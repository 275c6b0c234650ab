// Conjugate gradients on the translation normal equations  J^T J x = J^T t~
// (replaces `cg(incidence_r.T @ incidence_r, incidence_r.T @ t_tilde)`, vican/bipgo.py:476-478).
//
// The reference returns a TRUNCATED iterate of scipy.sparse.linalg.cg (rtol 1e-5), and on graphs
// with a few high-degree nodes (object calibration: 24 markers x ~1200 frames) that iterate is
// chaotic in the rounding of the mat-vec: the same recurrences with the row sums of the 24 marker
// rows taken in another order land 1e-6 .. 5e-6 away (measured with scipy alone).
// Parity therefore needs scipy's ARITHMETIC, not only its recurrences:
//
//   * scipy multiplies with the explicit CSR product J^T J (sorted column indices): row i is the
//     left-to-right sum over ascending unknown index of A_ij p_j, products rounded before they are
//     added (no FMA).  A_ij = -w_ct (w_ct = sum of k_t^2 over the pair's detections), A_ii = the
//     weighted degree, both never formed here: the mat-vec below evaluates EXACTLY that sum, one
//     sequential chain per (row, coordinate), with the diagonal term inserted at its sorted position.
//   * x += alpha p, r -= alpha q, p = beta p + r are evaluated with separately rounded products.
//   * the dot products are fixed-order two-stage reductions (deterministic, bitwise reproducible
//     run to run; their order differs from BLAS ddot, which moves the result by <= 5e-8).
//
// Data layout: both sides of the bipartite Laplacian are kept in a sliced-ELL layout (slices of 8
// rows, chunks of 4 columns): one chunk = 32 slots = one slot per lane, lane 4 j + s holds column
// 4 q + s of row j of the slice, so a warp streams (index, weight) with fully coalesced loads,
// gathers the 32 far-endpoint rows with 256-bit loads, and lanes (j, d < 3) run the 8 x 3
// sequential chains of the slice out of a small shared-memory stage.  Rows of a camera are long
// (5000 edges at cfg4, 1200 at cfg2): a whole CTA feeds the chains of one camera slice (8 chunks
// per step, double buffered); time rows are short: one warp per slice.
// Algorithmic bytes per iteration: 12 B per stored slot and side + 240 B per node (vectors).
#pragma once
#include <cub/cub.cuh>

#include "../../include/vican_b200.h"
#include "common.cuh"
#include "passes.cuh"
#include "peer.cuh"
#include "rotation.cuh"

namespace vb {

constexpr int SELL_ROWS = 8;
constexpr int CG_THREADS = 256;
constexpr int CG_WARPS = CG_THREADS / 32;
constexpr int CG_TIME_U = 4;        // chunks per step of a time-role warp
constexpr int CG_MAX_BLOCKS = 8192; // partial-table rows

inline int64_t sell_slices(int64_t n_rows) { return (n_rows + SELL_ROWS - 1) / SELL_ROWS; }

// ------------------------------------------------------------------------------ layout build
__global__ void sell_row_len_time_kernel(const int* __restrict__ rowptr, int64_t n_rows, int* __restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows) len[r] = rowptr[r + 1] - rowptr[r];
}

__global__ void sell_row_len_cam_kernel(const int* __restrict__ segptr, int64_t n_win, int64_t n_c, int* __restrict__ len) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    int s = 0;
    for (int64_t w = 0; w < n_win; ++w) s += segptr[w * n_c + c + 1] - segptr[w * n_c + c];
    len[c] = s;
}

// chunks of a slice = ceil(max row length / 4); entry n_slices is the scan's sentinel
__global__ void sell_count_kernel(const int* __restrict__ len, int64_t n_rows, int64_t n_slices, int* __restrict__ nchunk) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_slices) return;
    int m = 0;
    if (s < n_slices)
        for (int j = 0; j < SELL_ROWS; ++j) {
            const int64_t r = SELL_ROWS * s + j;
            if (r < n_rows) m = max(m, len[r]);
        }
    nchunk[s] = (m + 3) >> 2;
}

// warp per slice; padding slots carry idx = -1, w = 0
__global__ void sell_fill_time_kernel(const int* __restrict__ rowptr, const int* __restrict__ t_cam, const double* __restrict__ t_w,
                                      int64_t n_rows, int64_t n_slices, const int* __restrict__ ptr, int* __restrict__ idx,
                                      double* __restrict__ w) {
    const int lane = threadIdx.x & 31, j = lane >> 2, sub = lane & 3;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t slice = warp0; slice < n_slices; slice += nwarps) {
        const int64_t row = SELL_ROWS * slice + j;
        int s = 0, e = 0;
        if (row < n_rows) { s = rowptr[row]; e = rowptr[row + 1]; }
        const int c0 = ptr[slice], c1 = ptr[slice + 1];
        for (int q = c0; q < c1; ++q) {
            const int src = s + 4 * (q - c0) + sub;
            const bool ok = src < e;
            const int64_t slot = 32 * (int64_t)q + lane;
            idx[slot] = ok ? t_cam[src] : -1;
            w[slot] = ok ? t_w[src] : 0.0;
        }
    }
}

// camera rows: the camera-pass order keeps camera c's edges in n_win runs (window, c); their concatenation is
// the camera's edge list in ascending time order.  cam_prefix[w * n_c + c] = edges of camera c in windows < w.
__global__ void sell_cam_prefix_kernel(const int* __restrict__ segptr, int64_t n_win, int64_t n_c, int* __restrict__ prefix) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    int s = 0;
    for (int64_t w = 0; w < n_win; ++w) {
        prefix[w * n_c + c] = s;
        s += segptr[w * n_c + c + 1] - segptr[w * n_c + c];
    }
}

// One warp per (slice, range of SELL_FILL_CHUNKS chunks): lane (j, sub) locates element k = 4 q + sub of row
// 8 slice + j in the row's runs (binary search in the prefix table at the start of the range, then a monotone
// walk) and the warp writes whole chunks (coalesced 128 / 256-byte stores); padding: idx -1, w 0.
constexpr int SELL_FILL_CHUNKS = 64;
__global__ void sell_fill_cam_kernel(const int* __restrict__ segptr, const int* __restrict__ prefix, int64_t n_win, int64_t n_c,
                                     const int* __restrict__ c_time, const double* __restrict__ c_w, int64_t n_slices,
                                     const int* __restrict__ ptr, const int* __restrict__ item_ptr, int64_t n_items,
                                     int* __restrict__ idx, double* __restrict__ w) {
    const int lane = threadIdx.x & 31, j = lane >> 2, sub = lane & 3;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t item = warp0; item < n_items; item += nwarps) {
        // slice of this item: item_ptr[s] = first item of slice s (ascending): binary search
        int64_t lo = 0, hi = n_slices;
        while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (item_ptr[mid] <= item) lo = mid; else hi = mid; }
        const int64_t slice = lo;
        const int c0 = ptr[slice], c1 = ptr[slice + 1];
        const int qa = (int)(item - item_ptr[slice]) * SELL_FILL_CHUNKS;
        const int qb = (qa + SELL_FILL_CHUNKS < c1 - c0) ? qa + SELL_FILL_CHUNKS : c1 - c0;
        const int64_t c = SELL_ROWS * slice + j;
        // window holding element 4 qa + sub of camera c
        int64_t win = 0;
        int kbase = 0, rs = 0, re = 0;
        bool ok = c < n_c;
        if (ok) {
            const int k = 4 * qa + sub;
            int64_t a = 0, b2 = n_win;
            while (b2 - a > 1) { const int64_t mid = (a + b2) >> 1; if (prefix[mid * n_c + c] <= k) a = mid; else b2 = mid; }
            win = a; kbase = prefix[win * n_c + c]; rs = segptr[win * n_c + c]; re = segptr[win * n_c + c + 1];
        }
        for (int q = qa; q < qb; ++q) {
            const int k = 4 * q + sub;
            while (ok && k >= kbase + (re - rs)) {
                kbase += re - rs;
                if (++win >= n_win) { ok = false; break; }
                rs = segptr[win * n_c + c]; re = segptr[win * n_c + c + 1];
            }
            const int64_t slot = 32 * (int64_t)(c0 + q) + lane;
            const int src = rs + (k - kbase);
            idx[slot] = ok ? c_time[src] : -1;
            w[slot] = ok ? c_w[src] : 0.0;
        }
    }
}

// items per slice (ceil(chunks / SELL_FILL_CHUNKS)); entry n_slices is the scan's sentinel
__global__ void sell_items_kernel(const int* __restrict__ ptr, int64_t n_slices, int* __restrict__ cnt) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_slices) return;
    cnt[s] = s < n_slices ? (ptr[s + 1] - ptr[s] + SELL_FILL_CHUNKS - 1) / SELL_FILL_CHUNKS : 0;
}

struct SellWork {
    int *len_t, *len_c, *cnt_t, *cnt_c, *prefix;
    void* cub_tmp;
    size_t cub_bytes;
    int64_t bytes;
};

inline SellWork carve_sell(void* base, int64_t n_c, int64_t n_t, int64_t n_win) {
    SellWork w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void* r = p + off;
        off += (bytes + 255) & ~(int64_t)255;
        return r;
    };
    w.len_t = (int*)take(4 * (n_t + 1)); w.len_c = (int*)take(4 * (n_c + 1));
    w.cnt_t = (int*)take(4 * (sell_slices(n_t) + 2)); w.cnt_c = (int*)take(4 * (sell_slices(n_c) + 2));
    w.prefix = (int*)take(4 * (n_win * n_c + 1));
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)(sell_slices(n_t > n_c ? n_t : n_c) + 2));
    w.cub_bytes = b + 256;
    w.cub_tmp = take((int64_t)w.cub_bytes);
    w.bytes = off;
    return w;
}

inline int sell_count(const vb_graph* g, int32_t* st_ptr, int32_t* sc_ptr, int64_t* h_chunks_t, int64_t* h_chunks_c,
                      void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    SellWork w = carve_sell(workspace, n_c, n_t, g->n_windows);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    const int64_t ns_t = sell_slices(n_t), ns_c = sell_slices(n_c);
    int tot_t = 0, tot_c = 0;
    if (pinned_scalars() == nullptr) return -(int)cudaErrorMemoryAllocation;
    if (n_t > 0) {
        sell_row_len_time_kernel<<<(int)((n_t + 255) / 256), 256, 0, st>>>(g->t_rowptr, n_t, w.len_t);
        sell_count_kernel<<<(int)((ns_t + 256) / 256), 256, 0, st>>>(w.len_t, n_t, ns_t, w.cnt_t);
        size_t tb = w.cub_bytes;
        VB_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, (const int*)w.cnt_t, st_ptr, (int)(ns_t + 1), st));
        VB_CHECK(cudaMemcpyAsync(pinned_scalars(), st_ptr + ns_t, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    sell_row_len_cam_kernel<<<(int)((n_c + 255) / 256), 256, 0, st>>>(g->c_segptr, g->n_windows, n_c, w.len_c);
    sell_count_kernel<<<(int)((ns_c + 256) / 256), 256, 0, st>>>(w.len_c, n_c, ns_c, w.cnt_c);
    {
        size_t tb = w.cub_bytes;
        VB_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, (const int*)w.cnt_c, sc_ptr, (int)(ns_c + 1), st));
    }
    VB_CHECK(cudaMemcpyAsync(pinned_scalars() + 1, sc_ptr + ns_c, sizeof(int), cudaMemcpyDeviceToHost, st));
    VB_KERNEL_CHECK();
    VB_CHECK(cudaStreamSynchronize(st));
    if (n_t > 0) tot_t = *(const int*)pinned_scalars();
    tot_c = *(const int*)(pinned_scalars() + 1);
    count_launches(n_t > 0 ? 4 : 2);
    *h_chunks_t = tot_t;
    *h_chunks_c = tot_c;
    return 0;
}

inline int sell_fill(const vb_graph* g, const int32_t* st_ptr, int32_t* st_idx, double* st_w, const int32_t* sc_ptr,
                     int32_t* sc_idx, double* sc_w, int64_t chunks_c, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    SellWork w = carve_sell(workspace, n_c, n_t, g->n_windows);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    const int64_t ns_t = sell_slices(n_t), ns_c = sell_slices(n_c);
    const int cap = sm_count() * 8;
    if (ns_t > 0) {
        int grid = (int)((ns_t + CG_WARPS - 1) / CG_WARPS);
        sell_fill_time_kernel<<<grid < cap ? grid : cap, CG_THREADS, 0, st>>>(g->t_rowptr, g->t_cam, g->t_w, n_t, ns_t, st_ptr, st_idx, st_w);
    }
    {
        sell_cam_prefix_kernel<<<(int)((n_c + 255) / 256), 256, 0, st>>>(g->c_segptr, g->n_windows, n_c, w.prefix);
        sell_items_kernel<<<(int)((ns_c + 256) / 256), 256, 0, st>>>(sc_ptr, ns_c, w.cnt_c);
        size_t tb = w.cub_bytes;
        VB_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, (const int*)w.cnt_c, w.len_c, (int)(ns_c + 1), st));
        // upper bound of the item count without a read-back: every slice has < chunks / 64 + 1 items
        const int64_t items_max = chunks_c / SELL_FILL_CHUNKS + ns_c;
        int grid = (int)((items_max + CG_WARPS - 1) / CG_WARPS);
        sell_fill_cam_kernel<<<grid < cap ? grid : cap, CG_THREADS, 0, st>>>(g->c_segptr, w.prefix, g->n_windows, n_c, g->c_time, g->c_w, ns_c,
                                                                             sc_ptr, w.len_c, items_max, sc_idx, sc_w);
    }
    VB_KERNEL_CHECK();
    count_launches(ns_t > 0 ? 4 : 3);
    return 0;
}

// ------------------------------------------------------------------------------------- CG
enum { CG_RHO = 0, CG_BETA, CG_ALPHA, CG_DONE, CG_ITERS, CG_ATOL, CG_BN2, CG_RZ_C, CG_RZ_T, CG_RR_C, CG_RR_T, CG_PQ_C,
       CG_PQ_T, CG_NSCAL = 16 };
constexpr int CG_PQ_SLOT = 4;   // sharded runs: q_c[3 n_c + CG_PQ_SLOT .. + 1] carry the two halves of p . q through the collective
                                // (slots 0..2 behind the camera vector are the r.r / r.z pack)

struct CgWork {
    double *r_c, *p_c, *q_c, *dg_c;   // p padded [n][4]; q_c has 8 pack slots behind it
    double *r_t, *p_t, *q_t, *dg_t;
    int *ins_c, *ins_t;
    double *sc, *tab;
    unsigned* ticket;
    int64_t bytes;
};

inline CgWork carve_cg(void* base, int64_t n_c, int64_t n_t) {
    CgWork w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t nd) {
        double* r = (double*)(p + off);
        off += align256(nd * (int64_t)sizeof(double));
        return r;
    };
    w.r_c = take(3 * n_c); w.p_c = take(4 * n_c); w.q_c = take(3 * n_c + 8); w.dg_c = take(n_c);
    w.r_t = take(3 * n_t); w.p_t = take(4 * n_t); w.q_t = take(3 * n_t); w.dg_t = take(n_t);
    w.ins_c = (int*)take((n_c + 1) / 2 + 1); w.ins_t = (int*)take((n_t + 1) / 2 + 1);
    w.sc = take(CG_NSCAL);
    w.tab = take(2 * (int64_t)CG_MAX_BLOCKS);
    w.ticket = (unsigned*)take(4);
    w.bytes = off;
    return w;
}

// scipy cg, top of an iteration: `if norm(r) < atol: return` ; rho = r.z ; beta = rho / rho_prev
__device__ __forceinline__ void cg_top(double* sc, double rtol, int first) {
    const double rr = sc[CG_RR_C] + sc[CG_RR_T];
    const double rz = sc[CG_RZ_C] + sc[CG_RZ_T];
    if (first) {
        sc[CG_BN2] = rr;
        sc[CG_ATOL] = rtol * sqrt(rr);
        if (rr == 0.0) sc[CG_DONE] = 1.0;   // scipy: b == 0 -> x = 0
    }
    if (sc[CG_DONE] != 0.0) return;
    if (sqrt(rr) < sc[CG_ATOL]) { sc[CG_DONE] = 1.0; return; }
    sc[CG_BETA] = first ? 0.0 : rz / sc[CG_RHO];
    sc[CG_RHO] = rz;
    sc[CG_ITERS] += 1.0;
}

// Block sum of N per-thread values into row blockIdx.x of the partial table, then a ticket: returns
// true in every thread of the LAST block to arrive (all rows are then visible to it).
template <int N, int WARPS = CG_WARPS>
__device__ __forceinline__ bool cg_block_partial(const double (&v)[N], double* tab, unsigned* ticket) {
    __shared__ double sm[N][WARPS];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) sm[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += sm[threadIdx.x][w];
        tab[(size_t)blockIdx.x * N + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// fixed-order sum of component n of the table rows [b0, b1), by a whole block; result in every thread
template <int N, int THREADS = CG_THREADS>
__device__ __forceinline__ double cg_table_sum(const double* tab, int n, int b0, int b1) {
    __shared__ double sm[THREADS / 32];
    __shared__ double tot;
    double s = 0.0;
    for (int b = b0 + (int)threadIdx.x; b < b1; b += THREADS) s += __ldcg(tab + (size_t)b * N + n);
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += sm[w];
        tot = t;
    }
    __syncthreads();
    return tot;
}

struct CgVec {
    const double *b_c, *b_t;        // right-hand side (init only)
    double *x_c, *x_t, *r_c, *r_t, *p_c, *p_t;
    const double *q_c, *q_t, *dg_c, *dg_t;
    int64_t n_c, n_t;
    int nb_c;                       // blocks [0, nb_c) own cameras, the rest time nodes
    int jacobi, multi;
    double rtol;
    double *sc, *tab;
    unsigned* ticket;
    double* pack;                   // multi: local time-side partials for the cross-rank sum
};

// which node a thread owns
__device__ __forceinline__ bool cg_node(const CgVec& a, int64_t& i, bool& is_cam) {
    is_cam = (int)blockIdx.x < a.nb_c;
    i = is_cam ? (int64_t)blockIdx.x * CG_THREADS + threadIdx.x : (int64_t)(blockIdx.x - a.nb_c) * CG_THREADS + threadIdx.x;
    return i < (is_cam ? a.n_c : a.n_t);
}

__device__ __forceinline__ void cg_finish_rr(const CgVec& a, bool last, int first) {
    if (!last) return;
    const double rz_c = cg_table_sum<2>(a.tab, 0, 0, a.nb_c), rr_c = cg_table_sum<2>(a.tab, 1, 0, a.nb_c);
    const double rz_t = cg_table_sum<2>(a.tab, 0, a.nb_c, gridDim.x), rr_t = cg_table_sum<2>(a.tab, 1, a.nb_c, gridDim.x);
    if (threadIdx.x == 0) {
        a.sc[CG_RZ_C] = rz_c; a.sc[CG_RR_C] = rr_c; a.sc[CG_RZ_T] = rz_t; a.sc[CG_RR_T] = rr_t;
        if (a.multi) { a.pack[0] = rr_t; a.pack[1] = rz_t; a.pack[2] = 0.0; }
        else cg_top(a.sc, a.rtol, first);
        *a.ticket = 0u;
    }
}

// x = 0, r = b, p = 0; r.r and r.z
__global__ void __launch_bounds__(CG_THREADS) cg_init_kernel(CgVec a) {
    int64_t i; bool cam;
    const bool ok = cg_node(a, i, cam);
    double v[2] = {0.0, 0.0};
    if (ok) {
        const double* b = cam ? a.b_c : a.b_t;
        double* x = cam ? a.x_c : a.x_t; double* r = cam ? a.r_c : a.r_t; double* p = cam ? a.p_c : a.p_t;
        const double d = a.jacobi ? 1.0 / (cam ? a.dg_c : a.dg_t)[i] : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double bv = b[3 * i + k];
            x[3 * i + k] = 0.0; r[3 * i + k] = bv; p[4 * i + k] = 0.0;
            v[0] += bv * bv * d; v[1] += bv * bv;
        }
        p[4 * i + 3] = 0.0;
    }
    cg_finish_rr(a, cg_block_partial<2>(v, a.tab, a.ticket), 1);
}

// p = beta p + z   (scipy: p *= beta; p += z -- product rounded before the sum)
__global__ void __launch_bounds__(CG_THREADS) cg_dir_kernel(CgVec a) {
    if (a.sc[CG_DONE] != 0.0) return;
    int64_t i; bool cam;
    if (!cg_node(a, i, cam)) return;
    const double beta = a.sc[CG_BETA];
    const double* r = cam ? a.r_c : a.r_t; double* p = cam ? a.p_c : a.p_t;
    const double d = a.jacobi ? 1.0 / (cam ? a.dg_c : a.dg_t)[i] : 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double z = a.jacobi ? r[3 * i + k] * d : r[3 * i + k];
        p[4 * i + k] = __dadd_rn(__dmul_rn(beta, p[4 * i + k]), z);
    }
}

// x += alpha p; r -= alpha q; r.r and r.z; the last block prepares the next iteration's scalars
__global__ void __launch_bounds__(CG_THREADS) cg_update_kernel(CgVec a) {
    if (a.sc[CG_DONE] != 0.0) return;
    int64_t i; bool cam;
    const bool ok = cg_node(a, i, cam);
    double v[2] = {0.0, 0.0};
    if (ok) {
        // sharded: both halves of p . q arrive summed over ranks behind the camera vector (CG_PQ_SLOT)
        const double alpha = a.multi ? a.sc[CG_RHO] / (a.q_c[3 * a.n_c + CG_PQ_SLOT + 1] + a.q_c[3 * a.n_c + CG_PQ_SLOT])
                                     : a.sc[CG_ALPHA];
        double* x = cam ? a.x_c : a.x_t; double* r = cam ? a.r_c : a.r_t;
        const double* p = cam ? a.p_c : a.p_t; const double* q = cam ? a.q_c : a.q_t;
        const double d = a.jacobi ? 1.0 / (cam ? a.dg_c : a.dg_t)[i] : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[3 * i + k] = __dadd_rn(x[3 * i + k], __dmul_rn(alpha, p[4 * i + k]));
            const double rv = __dsub_rn(r[3 * i + k], __dmul_rn(alpha, q[3 * i + k]));
            r[3 * i + k] = rv;
            v[0] += rv * rv * d; v[1] += rv * rv;
        }
    }
    cg_finish_rr(a, cg_block_partial<2>(v, a.tab, a.ticket), 0);
}

// multi-rank: scalars after the cross-rank sum of the time-side partials
__global__ void cg_top_multi_kernel(double* sc, const double* pack, double rtol, int first) {
    if (!first && sc[CG_DONE] != 0.0) return;
    sc[CG_RR_T] = pack[0]; sc[CG_RZ_T] = pack[1];
    cg_top(sc, rtol, first);
}

// ------------------------------------------------------------------------------- mat-vec
// One side of the bipartite Laplacian in the sliced-ELL layout, as seen by the mat-vec.
struct SellSide {
    const int *ptr, *idx; const double* w; int64_t n_rows, n_slices;
    const double *p_other, *p_self;   // padded [n][4]: gather source (far endpoint) and the rows' own entries
    const double* dg;                 // weighted degrees (the diagonal of J^T J)
    const int* ins;                   // position of the diagonal term in the row's sorted order (nullptr: ins_default)
    int ins_default, add_diag;
    double* q;                        // [n][3]; diag mode: [n] row sums of the weights
};

struct CgMv {
    SellSide cam, time;
    double *sc, *tab;
    unsigned* ticket;
    int warps_cam;                    // warps [0, warps_cam) stream camera slices, the rest time slices
    int multi;                        // q_c is a partial sum over ranks: the halves of p . q ride with it (CG_PQ_SLOT)
};

constexpr int CG_MV_THREADS = 128;
constexpr int CG_MV_WARPS = CG_MV_THREADS / 32;
constexpr int CG_U = 4;               // chunks (of 32 slots) per pipeline step of a warp
constexpr int CG_MV_CTAS = 4;         // resident CTAs per SM the register budget is set for

__device__ __forceinline__ int ld_stream_i(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_d(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// A warp streams the contiguous chunk range of its slices [s0, s1).  Lane (j, d) owns the chain of coordinate d
// of row j of the current slice and does everything for it itself: it reads its row's four (index, weight)
// pairs of a chunk with one 128-bit and one 256-bit load (the three lanes of a row read the same addresses:
// one request), gathers component d of the four far-endpoint rows (the three lanes of a row hit one 32-byte
// sector), and advances the chain in registers -- no shared memory, no cross-lane traffic.  Two-stage software
// pipeline: (index, weight) of step s+1 are in flight while step s is gathered and chained; the other warps of
// the SM cover the gather latency.  L1 wavefronts per chunk: 32 (gathers) + 3, the floor for this access pattern.
// Slice boundaries inside a step are handled in the chain loop (rows finish, the next slice's rows start).
template <bool DIAGMODE>
__device__ __forceinline__ void cg_stream_rows(const SellSide& S, int64_t s0, int64_t s1, double& pq) {
    if (s0 >= s1) return;
    const int lane = threadIdx.x & 31, j = lane >> 2, d = lane & 3;
    const int cbeg = __ldg(S.ptr + s0), cend = __ldg(S.ptr + s1);
    int4 idx_nx[CG_U]; double w_nx[CG_U][4];
    auto loadA = [&](int q0) {
#pragma unroll
        for (int u = 0; u < CG_U; ++u) {
            idx_nx[u] = make_int4(-1, -1, -1, -1);
            w_nx[u][0] = w_nx[u][1] = w_nx[u][2] = w_nx[u][3] = 0.0;
            if (q0 + u < cend) {
                const int64_t base = 32 * (int64_t)(q0 + u) + 4 * j;
                asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(idx_nx[u].x), "=r"(idx_nx[u].y), "=r"(idx_nx[u].z), "=r"(idx_nx[u].w) : "l"(S.idx + base));
                asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                             : "=d"(w_nx[u][0]), "=d"(w_nx[u][1]), "=d"(w_nx[u][2]), "=d"(w_nx[u][3]) : "l"(S.w + base));
            }
        }
    };
    // chain state of the current slice (+ the next slice's row constants, fetched one slice ahead)
    int64_t slice = s0;
    int slice_end = __ldg(S.ptr + s0 + 1);
    double acc = 0.0, diag = 0.0, pself = 0.0, n_pself = 0.0, n_dg = 0.0;
    int ins = 0x7fffffff, n_ins = 0x7fffffff, k = 0;
    bool pending = false, active = false;
    auto fetch_row = [&](int64_t sl) {   // constants of slice sl into the n_* registers
        const int64_t row = SELL_ROWS * sl + j;
        n_pself = 0.0; n_dg = 0.0; n_ins = S.ins_default;
        if (!DIAGMODE && sl < s1 && d < 3 && row < S.n_rows) {
            n_pself = S.p_self[4 * row + d];
            if (S.add_diag) { n_dg = S.dg[row]; if (S.ins) n_ins = S.ins[row]; }
        }
    };
    auto start_row = [&](int64_t sl) {
        const int64_t row = SELL_ROWS * sl + j;
        active = (d < 3) && (row < S.n_rows);
        acc = 0.0; k = 0; pself = n_pself; pending = false; ins = 0x7fffffff;
        if (!DIAGMODE && active && S.add_diag) { diag = __dmul_rn(n_dg, pself); pending = true; ins = n_ins; }
        fetch_row(sl + 1);
    };
    auto finish_row = [&](int64_t sl) {
        if (!active) return;
        const int64_t row = SELL_ROWS * sl + j;
        if (pending) acc = __dadd_rn(acc, diag);
        if (DIAGMODE) { if (d == 0) S.q[row] = acc; }
        else {
            S.q[3 * row + d] = acc;
            pq += pself * acc;
        }
    };
    fetch_row(s0);
    start_row(s0);
    loadA(cbeg);
    const int dd = d < 3 ? d : 0;
    const uint64_t keep = policy_evict_last();   // the gathered vectors are re-read by many rows
    for (int q0 = cbeg; q0 < cend; q0 += CG_U) {
        // gathers of this step (component d of the four far-endpoint rows of every chunk), then the next step's
        // (index, weight) stream; the products are rounded before they are added (scipy: no FMA)
        double pr[CG_U][4];
#pragma unroll
        for (int u = 0; u < CG_U; ++u) {
            const int id[4] = {idx_nx[u].x, idx_nx[u].y, idx_nx[u].z, idx_nx[u].w};
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
                double pv = DIAGMODE ? 1.0 : 0.0;
                if (!DIAGMODE && id[sub] >= 0) pv = ld_keep(S.p_other + 4 * (int64_t)id[sub] + dd, keep);
                pr[u][sub] = pv;
            }
        }
#pragma unroll
        for (int u = 0; u < CG_U; ++u)
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) pr[u][sub] = __dmul_rn(w_nx[u][sub], pr[u][sub]);
        loadA(q0 + CG_U);
#pragma unroll
        for (int u = 0; u < CG_U; ++u) {
            const int c = q0 + u;
            if (c < cend) {        // warp-uniform
                while (c == slice_end) {   // the previous slice's rows are complete (warp-uniform)
                    finish_row(slice);
                    ++slice;
                    slice_end = __ldg(S.ptr + slice + 1);
                    start_row(slice);
                }
                if (pending && ins < k + 4) {
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        if (pending && k == ins) { acc = __dadd_rn(acc, diag); pending = false; }
                        acc = DIAGMODE ? __dadd_rn(acc, pr[u][sub]) : __dsub_rn(acc, pr[u][sub]);
                        ++k;
                    }
                } else {
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) acc = DIAGMODE ? __dadd_rn(acc, pr[u][sub]) : __dsub_rn(acc, pr[u][sub]);
                    k += 4;
                }
            }
        }
    }
    // the last slice with chunks, and any trailing slices without (rows with no local edges)
    for (;;) {
        finish_row(slice);
        if (++slice >= s1) break;
        start_row(slice);
    }
}

// q = (J^T J) p in scipy's CSR row order, or (DIAGMODE) the weighted degrees by the same sequential sums
template <bool DIAGMODE>
__global__ void __launch_bounds__(CG_MV_THREADS, CG_MV_CTAS) cg_matvec_kernel(CgMv a) {
    if (!DIAGMODE && a.sc[CG_DONE] != 0.0) return;
    const int wv = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * CG_MV_WARPS + wv;
    const int64_t nw = (int64_t)gridDim.x * CG_MV_WARPS;
    double pq = 0.0;
    if (gw < a.warps_cam) {
        const int64_t per = (a.cam.n_slices + a.warps_cam - 1) / a.warps_cam;
        const int64_t s0 = gw * per, s1 = (s0 + per < a.cam.n_slices) ? s0 + per : a.cam.n_slices;
        cg_stream_rows<DIAGMODE>(a.cam, s0, s1, pq);
    } else {
        const int64_t wt = nw - a.warps_cam, me = gw - a.warps_cam;
        const int64_t per = (a.time.n_slices + wt - 1) / (wt > 0 ? wt : 1);
        const int64_t s0 = me * per, s1 = (s0 + per < a.time.n_slices) ? s0 + per : a.time.n_slices;
        cg_stream_rows<DIAGMODE>(a.time, s0, s1, pq);
    }
    if (DIAGMODE) return;
    // p . q: camera warps own whole CTAs [0, warps_cam / 4) (the host rounds warps_cam to CTAs), so the two parts
    // are summed separately and in a fixed order
    const double v[1] = {pq};
    if (cg_block_partial<1, CG_MV_WARPS>(v, a.tab, a.ticket)) {
        const int nb_cam = a.warps_cam / CG_MV_WARPS;
        const double pq_c = cg_table_sum<1, CG_MV_THREADS>(a.tab, 0, 0, nb_cam);
        const double pq_t = cg_table_sum<1, CG_MV_THREADS>(a.tab, 0, nb_cam, gridDim.x);
        if (threadIdx.x == 0) {
            a.sc[CG_PQ_C] = pq_c; a.sc[CG_PQ_T] = pq_t;
            if (a.multi) {
                // q_c is this rank's partial sum and the dot product is linear in it: both halves ride with the
                // camera vector through the collective, cg_update_kernel divides
                a.cam.q[3 * a.cam.n_rows + CG_PQ_SLOT] = pq_t; a.cam.q[3 * a.cam.n_rows + CG_PQ_SLOT + 1] = pq_c;
            } else {
                a.sc[CG_ALPHA] = a.sc[CG_RHO] / (pq_c + pq_t);
            }
            *a.ticket = 0u;
        }
    }
}

// position of the diagonal entry in a row's ascending-unknown-index order = number of neighbours
// whose unknown index is smaller than the row's own (warp per slice)
__global__ void cg_ins_kernel(const int* __restrict__ ptr, const int* __restrict__ idx, int64_t n_rows, int64_t n_slices,
                              const int* __restrict__ unk_self, const int* __restrict__ unk_other, int* __restrict__ ins) {
    const int lane = threadIdx.x & 31, j = lane >> 2;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t slice = warp0; slice < n_slices; slice += nwarps) {
        const int64_t row = SELL_ROWS * slice + j;
        const int me = row < n_rows ? unk_self[row] : 0;
        int cnt = 0;
        for (int q = ptr[slice]; q < ptr[slice + 1]; ++q) {
            const int o = idx[32 * (int64_t)q + lane];
            if (o >= 0 && unk_other[o] < me) ++cnt;
        }
        cnt += __shfl_xor_sync(FULL, cnt, 1);
        cnt += __shfl_xor_sync(FULL, cnt, 2);
        if ((lane & 3) == 0 && row < n_rows) ins[row] = cnt;
    }
}

inline int cg_occupancy_blocks() {
    static int n = 0;
    if (n == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_matvec_kernel<false>, CG_MV_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        n = per_sm * sm_count();
    }
    return n;
}

inline int trans_cg(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t, double rtol,
                    int64_t maxiter, int jacobi, const int32_t* unk_c, const int32_t* unk_t, int32_t* h_iters,
                    void* workspace, int64_t workspace_bytes, vb_allreduce_fn allreduce, void* actx, int owner,
                    cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    if (g->sc_ptr == nullptr || (n_t > 0 && g->st_ptr == nullptr)) return VB_STATUS_BAD_ARGUMENT;
    CgWork w = carve_cg(workspace, n_c, n_t);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    const int nb_c = (int)((n_c + CG_THREADS - 1) / CG_THREADS), nb_t = (int)((n_t + CG_THREADS - 1) / CG_THREADS);
    if (nb_c + nb_t > CG_MAX_BLOCKS) return VB_STATUS_BAD_ARGUMENT;
    VB_CHECK(cudaMemsetAsync(w.sc, 0, CG_NSCAL * sizeof(double), st));
    VB_CHECK(cudaMemsetAsync(w.ticket, 0, 4 * sizeof(unsigned), st));
    const int multi = allreduce != nullptr;
    const int64_t ns_t = sell_slices(n_t), ns_c = sell_slices(n_c);

    CgMv mv;
    mv.cam = SellSide{g->sc_ptr, g->sc_idx, g->sc_w, n_c, ns_c, w.p_t, w.p_c, w.dg_c, nullptr, 0, owner, w.q_c};
    mv.time = SellSide{g->st_ptr, g->st_idx, g->st_w, n_t, ns_t, w.p_c, w.p_t, w.dg_t, nullptr, 0x7fffffff, 1, w.q_t};
    mv.sc = w.sc; mv.tab = w.tab; mv.ticket = w.ticket; mv.multi = multi;
    // One resident wave of 4-warp CTAs.  Camera warps come first, in whole CTAs, one slice each where the
    // wave allows (a camera slice is one long serial chain per row: its time is its chain, so slices want to
    // run side by side); the remaining warps split the time slices evenly (contiguous ranges).
    const int cap = cg_occupancy_blocks();
    static const double cam_frac = getenv("VICAN_B200_CG_CAMFRAC") ? atof(getenv("VICAN_B200_CG_CAMFRAC")) : 0.55;
    int64_t cam_ctas = (ns_c + CG_MV_WARPS - 1) / CG_MV_WARPS;
    const int64_t cam_cap = (int64_t)(cap * cam_frac) > 1 ? (int64_t)(cap * cam_frac) : 1;
    if (cam_ctas > cam_cap) cam_ctas = cam_cap;
    int64_t time_ctas = (ns_t + CG_MV_WARPS - 1) / CG_MV_WARPS;
    if (time_ctas > cap - cam_ctas) time_ctas = cap - cam_ctas;
    if (time_ctas < 1 && ns_t > 0) time_ctas = 1;
    if (cam_ctas + time_ctas > CG_MAX_BLOCKS) return VB_STATUS_BAD_ARGUMENT;
    mv.warps_cam = (int)cam_ctas * CG_MV_WARPS;
    const int mv_grid = (int)(cam_ctas + time_ctas);

    if (unk_c != nullptr && unk_t != nullptr) {
        const int capw = sm_count() * 8;
        int gc = (int)((ns_c + CG_WARPS - 1) / CG_WARPS), gt = (int)((ns_t + CG_WARPS - 1) / CG_WARPS);
        cg_ins_kernel<<<gc < capw ? gc : capw, CG_THREADS, 0, st>>>(g->sc_ptr, g->sc_idx, n_c, ns_c, unk_c, unk_t, w.ins_c);
        if (ns_t > 0) cg_ins_kernel<<<gt < capw ? gt : capw, CG_THREADS, 0, st>>>(g->st_ptr, g->st_idx, n_t, ns_t, unk_t, unk_c, w.ins_t);
        mv.cam.ins = w.ins_c; mv.time.ins = w.ins_t;
    }
    {   // weighted degrees = diagonal of J^T J: the same sequential row sums, over the weights
        CgMv dm = mv;
        dm.cam.q = w.dg_c; dm.time.q = w.dg_t;
        cg_matvec_kernel<true><<<mv_grid, CG_MV_THREADS, 0, st>>>(dm);
        VB_KERNEL_CHECK();
        if (allreduce) { int rc = allreduce(actx, w.dg_c, n_c, (void*)st); if (rc) return rc; }
    }
    CgVec v;
    v.b_c = rhs_c; v.b_t = rhs_t; v.x_c = x_c; v.x_t = x_t; v.r_c = w.r_c; v.r_t = w.r_t; v.p_c = w.p_c; v.p_t = w.p_t;
    v.q_c = w.q_c; v.q_t = w.q_t; v.dg_c = w.dg_c; v.dg_t = w.dg_t; v.n_c = n_c; v.n_t = n_t; v.nb_c = nb_c;
    v.jacobi = jacobi; v.multi = multi; v.rtol = rtol; v.sc = w.sc; v.tab = w.tab; v.ticket = w.ticket;
    v.pack = w.q_c + 3 * n_c;
    const int vgrid = nb_c + nb_t;
    cg_init_kernel<<<vgrid, CG_THREADS, 0, st>>>(v);
    VB_KERNEL_CHECK();
    if (multi) {
        int rc = allreduce(actx, v.pack, 3, (void*)st);
        if (rc) return rc;
        cg_top_multi_kernel<<<1, 1, 0, st>>>(w.sc, v.pack, rtol, 1);
    }
    // The host polls the done flag one batch late: batch k+1 is enqueued before the flag of batch k is
    // read, so the GPU never idles on the host; iterations enqueued after convergence return at once.
    PinnedStatus& ps = pinned_state();
    int status = VB_STATUS_NOT_CONVERGED;
    // peer-memory collectives skip themselves once the done flag is up (identical on all ranks)
    PeerSkipScope skip_scope(allreduce == (vb_allreduce_fn)&vb_peer_allreduce ? (PeerCtx*)actx : nullptr, w.sc + CG_DONE);
    const int batch = 4;
    int64_t enq = 0;
    double* hs = ps.h;
    auto readback = [&](int b) -> int {
        const int slot = b % STATUS_SLOTS;
        VB_CHECK(cudaMemcpyAsync(ps.h + slot * SM_SIZE, w.sc, CG_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
        VB_CHECK(cudaEventRecord(ps.ev[slot], st));
        return 0;
    };
    { int rc = readback(0); if (rc) return rc; }
    for (int b = 1;; ++b) {
        for (int i = 0; i < batch && enq < maxiter; ++i, ++enq) {
            cg_dir_kernel<<<vgrid, CG_THREADS, 0, st>>>(v);
            cg_matvec_kernel<false><<<mv_grid, CG_MV_THREADS, 0, st>>>(mv);
            if (multi) {
                int rc = allreduce(actx, w.q_c, 3 * n_c + 8, (void*)st);
                if (rc) return rc;
            }
            cg_update_kernel<<<vgrid, CG_THREADS, 0, st>>>(v);
            VB_KERNEL_CHECK();
            if (multi) {
                int rc = allreduce(actx, v.pack, 3, (void*)st);
                if (rc) return rc;
                cg_top_multi_kernel<<<1, 1, 0, st>>>(w.sc, v.pack, rtol, 0);
            }
        }
        { int rc = readback(b); if (rc) return rc; }
        VB_CHECK(cudaEventSynchronize(ps.ev[(b - 1) % STATUS_SLOTS]));
        hs = ps.h + ((b - 1) % STATUS_SLOTS) * SM_SIZE;
        if (hs[CG_DONE] != 0.0) { status = VB_STATUS_OK; break; }
        if (enq >= maxiter) {
            VB_CHECK(cudaEventSynchronize(ps.ev[b % STATUS_SLOTS]));
            hs = ps.h + (b % STATUS_SLOTS) * SM_SIZE;
            if (hs[CG_DONE] != 0.0) status = VB_STATUS_OK;
            break;
        }
    }
    VB_CHECK(cudaStreamSynchronize(st));   // the speculative tail (no-ops) must not outlive the workspace
    // executed launches: degrees, init, 3 per executed iteration (+ the scalar kernels of a sharded run);
    // iterations enqueued after convergence return at their first instruction and are not counted
    count_launches(2 + (unk_c != nullptr ? 2 : 0) + (multi ? 1 : 0) + (long long)hs[CG_ITERS] * (multi ? 4 : 3));
    if (h_iters) *h_iters = (int32_t)hs[CG_ITERS];
    return status;
}

}  // namespace vb
