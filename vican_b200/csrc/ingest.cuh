// Edge ingestion: raw detections -> device-resident block-CSR (by time) + block-CSC (by camera).
//
// Replaces the Python dict loops and scipy COO->CSR conversions of vican/bipgo.py:203-276 and
// the node/edge indexing of :420-431.  The constraint fold blk = (k_r R_cm) R_m^T R_0 (:209-213)
// and the per-(camera, time) aggregation (:215-221) run on the device; aggregation order is
// the original detection order (stable sort), i.e. the reference's dict-insertion order.
// Key sorting uses CUB's device radix sort (index plumbing, not arithmetic).
#pragma once
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "../../include/vican_b200.h"
#include "common.cuh"

namespace vb {

// 1 where sorted position i starts a new (time, camera) pair, 0 at i = 0: the inclusive scan of these flags is
// the pair id of every position (fed to CUB's scan through a transform iterator: no flag array, no extra pass)
struct HeadFlagCT {
    const int* time; const int* cam;
    __host__ __device__ int operator()(int i) const { return (i > 0 && (time[i] != time[i - 1] || cam[i] != cam[i - 1])) ? 1 : 0; }
};
struct HeadFlagKey {
    const uint64_t* keys;
    __host__ __device__ int operator()(int i) const { return (i > 0 && keys[i] != keys[i - 1]) ? 1 : 0; }
};

// pair id one past the last pair that lies completely inside the first raw_end[k] sorted detections
__global__ void chunk_pairs_kernel(const int* __restrict__ raw_pair, const int64_t* __restrict__ raw_end, int n_chunks, int64_t n_raw,
                                   int64_t n_pairs, int64_t* __restrict__ pair_end) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chunks) return;
    const int64_t e = raw_end[k];
    pair_end[k] = e >= n_raw ? n_pairs : (int64_t)raw_pair[e];   // the pair of detection e may straddle the boundary: excluded
}

__global__ void check_sorted_ct_kernel(const int* __restrict__ time, const int* __restrict__ cam, int64_t n, int* __restrict__ unsorted_flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const int t0 = time[i], t1 = time[i + 1];
    if (t0 > t1 || (t0 == t1 && cam[i] > cam[i + 1])) *unsorted_flag = 1;
}

constexpr int ING_THREADS = 256;
inline int ing_grid(int64_t n) { return (int)((n + ING_THREADS - 1) / ING_THREADS); }

inline int key_bits(int64_t n_c, int64_t n_t) {
    unsigned long long m = (unsigned long long)n_c * (unsigned long long)n_t;
    int b = 1;
    while (b < 64 && (m >> b) != 0) ++b;
    return b;
}

__global__ void make_keys_kernel(const int* __restrict__ major, const int* __restrict__ minor, int64_t n_minor,
                                 uint64_t* __restrict__ keys, int* __restrict__ vals, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint64_t)major[i] * (uint64_t)n_minor + (uint64_t)minor[i];
    vals[i] = (int)i;
}

__global__ void pair_start_kernel(const int* __restrict__ raw_pair, int* __restrict__ pair_start, int64_t n_raw, int64_t n_pairs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_raw) return;
    const int p = raw_pair[i];
    if (i == 0 || raw_pair[i - 1] != p) pair_start[p] = (int)i;
    if (i == n_raw - 1) pair_start[n_pairs] = (int)n_raw;
}

// Per aggregated pair: its endpoints and its key in the camera-pass order, key = window(time) * n_c + cam
// (32-bit: n_windows * n_c is a few hundred thousand).  The pairs are already sorted by (time, camera) and the
// radix sort is stable, so sorting by this key orders them by (window, camera, time).  Tiles of all cameras
// that fall in the same time window are adjacent in the stream, so the W records gathered by concurrently
// running warps of the camera pass come from one window of W (L2 resident) instead of all of it.
__global__ void pair_keys_kernel(const int* __restrict__ cam, const int* __restrict__ time, const int* __restrict__ raw_perm,
                                 const int* __restrict__ pair_start, int64_t n_pairs, int64_t n_c, int64_t n_t, int64_t n_win,
                                 int* __restrict__ t_cam, int* __restrict__ t_time, uint32_t* __restrict__ keys,
                                 int* __restrict__ vals) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int64_t r0 = raw_perm[pair_start[p]];
    const int c = cam[r0], t = time[r0];
    t_cam[p] = c;
    t_time[p] = t;
    const uint64_t win = (uint64_t)t * (uint64_t)n_win / (uint64_t)n_t;
    keys[p] = (uint32_t)(win * (uint64_t)n_c + (uint64_t)c);
    vals[p] = (int)p;
}

// c_pos[c_order[i]] = i : where every time-sorted pair sits in the camera-pass order
__global__ void inverse_perm_kernel(const int* __restrict__ c_order, int* __restrict__ c_pos, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) c_pos[c_order[i]] = (int)i;
}

// Fold + aggregate (bipgo.py:209-221) writing BOTH copies of the blocks from the same registers: the
// time-sorted block-CSR (t_B, coalesced through a shared-memory transpose) and the camera-pass copy (c_B,
// stored transposed, one contiguous 72-byte record per pair at its camera-pass position), plus the
// per-pair weights in both orders.  One thread folds one pair, in original detection order.
constexpr int FOLD_THREADS = 256;
constexpr int FOLD_MAX_MARKERS_SMEM = 64;
__global__ void __launch_bounds__(FOLD_THREADS)
fold_both_kernel(const int* __restrict__ marker, const double* __restrict__ R, const double* __restrict__ k_r,
                 const double* __restrict__ k_t, const double* __restrict__ markerC, int round_f32,
                 const int* __restrict__ raw_perm, const int* __restrict__ pair_start, int64_t p_begin, int64_t n_pairs,
                 const int* __restrict__ t_time, const int* __restrict__ c_pos, double* __restrict__ t_B,
                 double* __restrict__ t_a, double* __restrict__ t_w, double* __restrict__ c_B, int* __restrict__ c_time,
                 double* __restrict__ c_w, int n_markers, int identity_perm) {
    // the chain of dependent global loads per pair (pair_start -> raw_perm -> detection -> marker constant) is what
    // bounds this kernel's latency: the marker constants sit in shared memory, sorted input needs no permutation
    __shared__ double sC[9 * FOLD_MAX_MARKERS_SMEM];
    const bool c_smem = n_markers <= FOLD_MAX_MARKERS_SMEM;
    if (c_smem)
        for (int i = threadIdx.x; i < 9 * n_markers; i += FOLD_THREADS) sC[i] = markerC[i];
    const double* mC = c_smem ? sC : markerC;
    __shared__ double sB[FOLD_THREADS * 9];
    __shared__ double sW[FOLD_THREADS];
    __shared__ int sPos[FOLD_THREADS], sTime[FOLD_THREADS];
    const int64_t p0 = p_begin + (int64_t)blockIdx.x * FOLD_THREADS;   // pairs [p_begin, n_pairs) of this launch
    const int64_t p = p0 + threadIdx.x;
    __syncthreads();
    if (p < n_pairs) {
        const int s = pair_start[p], e = pair_start[p + 1];
        double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double a = 0.0, w = 0.0;
        for (int pos = s; pos < e; ++pos) {
            const int64_t r = identity_perm ? pos : raw_perm[pos];
            const double kr = k_r[r], kt = k_t[r];
            double kR[9], Cm[9], blk[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double v = R[9 * r + i];
                kR[i] = round_f32 ? (double)((float)kr * (float)v) : kr * v;
                Cm[i] = mC[9 * (int64_t)marker[r] + i];
            }
            mm3(kR, Cm, blk);
#pragma unroll
            for (int i = 0; i < 9; ++i) B[i] += blk[i];
            a += kr;
            w += kt * kt;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) sB[9 * threadIdx.x + i] = B[i];
        t_a[p] = a;
        t_w[p] = w;
        sW[threadIdx.x] = w;
        sPos[threadIdx.x] = c_pos[p];
        sTime[threadIdx.x] = t_time[p];
    }
    __syncthreads();
    const int n_here = (int)((n_pairs - p0) < FOLD_THREADS ? (n_pairs - p0) : FOLD_THREADS);
    // time-sorted copy: the CTA's 256 x 9 doubles are contiguous in t_B
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int i = q * FOLD_THREADS + threadIdx.x;
        if (i < 9 * n_here) t_B[9 * p0 + i] = sB[i];
    }
    // camera-pass copy: 9 consecutive threads write one pair's record, transposed
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int i = q * FOLD_THREADS + threadIdx.x;
        const int pl = i / 9, k = i - 9 * pl;
        if (pl < n_here) {
            const int64_t dst = sPos[pl];
            c_B[9 * dst + k] = sB[9 * pl + 3 * (k % 3) + k / 3];
            if (k == 0) c_time[dst] = sTime[pl];
            if (k == 1) c_w[dst] = sW[pl];
        }
    }
}

__global__ void iota_kernel(int* __restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}

// ptr[v] = first position whose (sorted) node id is >= v, for v in [0, n_nodes]
__global__ void seg_ptr_kernel(const int* __restrict__ node_sorted, const int* __restrict__ perm, int* __restrict__ ptr,
                               int64_t n, int64_t n_nodes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cur = perm ? node_sorted[perm[i]] : node_sorted[i];
    const int prev = (i == 0) ? -1 : (perm ? node_sorted[perm[i - 1]] : node_sorted[i - 1]);
    for (int v = prev + 1; v <= cur; ++v) ptr[v] = (int)i;
    if (i == n - 1)
        for (int64_t v = cur + 1; v <= n_nodes; ++v) ptr[v] = (int)n;
}

// warp per node: deg[v] = sum of a over its segment (optionally through a permutation)
__global__ void seg_sum_kernel(const int* __restrict__ ptr, const int* __restrict__ perm, const double* __restrict__ a,
                               double* __restrict__ deg, int64_t n_nodes) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_nodes) return;
    const int s = ptr[warp], e = ptr[warp + 1];
    double acc = 0.0;
    for (int i = s + lane; i < e; i += 32) acc += a[perm ? perm[i] : i];
    acc = warp_sum(acc);
    if (lane == 0) deg[warp] = acc;
}

// entry n_seg is the sentinel of the exclusive scan (tile_off[n_seg] = number of tiles)
__global__ void tile_count_kernel(const int* __restrict__ colptr, int* __restrict__ cnt, int64_t n_seg, int tile_len) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_seg) return;
    const int len = c < n_seg ? colptr[c + 1] - colptr[c] : 0;
    cnt[c] = (len + tile_len - 1) / tile_len;
}

// tiles are contiguous: tile_start doubles as a segment pointer array (sentinel tile_start[n_tiles] = E)
__global__ void tile_fill_kernel(const int* __restrict__ colptr, const int* __restrict__ off, int* __restrict__ tile_cam,
                                 int* __restrict__ tile_start, int64_t n_seg, int64_t n_c, int tile_len) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_seg) return;
    const int s = colptr[c], e = colptr[c + 1];
    int t = off[c];
    for (int b = s; b < e; b += tile_len, ++t) {
        tile_cam[t] = (int)(c % n_c);
        tile_start[t] = b;
    }
    if (c == n_seg - 1) tile_start[off[n_seg]] = e;
}

// warp per camera: out[c] = sum over the camera's (window, camera) runs of a[order[i]] -- the
// camera-pass order keeps every camera's edges in n_win contiguous runs, so per-camera reductions
// need no camera-major copy (fixed summation order: deterministic)
__global__ void cam_runs_sum_kernel(const int* __restrict__ segptr, int64_t n_win, int64_t n_c, const int* __restrict__ order,
                                    const double* __restrict__ a, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_c) return;
    double acc = 0.0;
    for (int64_t wdw = 0; wdw < n_win; ++wdw) {
        const int s = segptr[wdw * n_c + c], e = segptr[wdw * n_c + c + 1];
        for (int i = s + lane; i < e; i += 32) acc += a[order ? order[i] : i];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[c] = acc;
}

// Connected components of the bipartite graph (cameras 0 .. n_c-1, time nodes n_c ..): min-label hooking over
// the aggregated edges + pointer jumping until nothing changes (a handful of rounds: the graph has a small
// diameter).  The reference's early exit (bipgo.py:283) can only fire when there is more than one component.
__global__ void cc_hook_kernel(const int* __restrict__ t_cam, const int* __restrict__ t_time, int64_t n_edges, int n_c,
                               int* __restrict__ label, int* __restrict__ changed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int u = t_cam[e], v = n_c + t_time[e];
    const int lu = label[u], lv = label[v];
    if (lu < lv) { atomicMin(label + lv, lu); atomicMin(label + v, lu); *changed = 1; }
    else if (lv < lu) { atomicMin(label + lu, lv); atomicMin(label + u, lv); *changed = 1; }
}
__global__ void cc_jump_kernel(int* __restrict__ label, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int l = label[i];
    while (label[l] != l) l = label[l];
    label[i] = l;
}
__global__ void cc_count_kernel(const int* __restrict__ label, int64_t n, int* __restrict__ count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && label[i] == (int)i) atomicAdd(count, 1);
}

struct IngestWork {
    uint64_t *keys_a, *keys_b;
    int *vals_a, *tmp_a, *tmp_b, *tmp_c, *tmp_d;
    void* cub_tmp;
    size_t cub_bytes;
    int64_t bytes;
};

inline size_t cub_temp_bytes(int64_t n) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)n, 0, 64);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, (int)n);
    return (a > b ? a : b) + 1024;
}

inline IngestWork carve_ingest(void* base, int64_t n) {
    IngestWork w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void* r = p + off;
        off += (bytes + 255) & ~(int64_t)255;
        return r;
    };
    w.keys_a = (uint64_t*)take(8 * n);
    w.keys_b = (uint64_t*)take(8 * n);
    w.vals_a = (int*)take(4 * n);
    w.tmp_a = (int*)take(4 * (n + 1));
    w.tmp_b = (int*)take(4 * (n + 1));
    w.tmp_c = (int*)take(4 * (n + 1));
    w.tmp_d = (int*)take(4 * (n + 1));
    w.cub_bytes = cub_temp_bytes(n);
    w.cub_tmp = take((int64_t)w.cub_bytes);
    w.bytes = off;
    return w;
}

inline int ingest_sort(const int* cam, const int* time, int64_t n_raw, int64_t n_c, int64_t n_t, int* raw_perm,
                       int* raw_pair, int64_t* h_n_pairs, int32_t* h_sorted, void* workspace, int64_t workspace_bytes,
                       cudaStream_t st) {
    if (n_raw <= 0) return VB_STATUS_BAD_ARGUMENT;
    IngestWork w = carve_ingest(workspace, n_raw);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    // adaptive: detections that already arrive ordered by (time, camera) -- the usual layout of a
    // recording -- skip the key build and the radix sort altogether
    VB_CHECK(cudaMemsetAsync(w.tmp_a, 0, sizeof(int), st));
    check_sorted_ct_kernel<<<ing_grid(n_raw), ING_THREADS, 0, st>>>(time, cam, n_raw, w.tmp_a);
    int unsorted = 0;
    VB_CHECK(read_back(&unsorted, w.tmp_a, sizeof(int), st));
    size_t tb = w.cub_bytes;
    thrust::counting_iterator<int> iota(0);
    if (unsorted) {
        make_keys_kernel<<<ing_grid(n_raw), ING_THREADS, 0, st>>>(time, cam, n_c, w.keys_a, w.vals_a, n_raw);
        VB_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, (const uint64_t*)w.keys_a, w.keys_b, (const int*)w.vals_a,
                                                 raw_perm, (int)n_raw, 0, key_bits(n_c, n_t), st));
        tb = w.cub_bytes;
        auto flags = thrust::make_transform_iterator(iota, HeadFlagKey{w.keys_b});
        VB_CHECK(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, flags, raw_pair, (int)n_raw, st));
    } else {
        iota_kernel<<<ing_grid(n_raw), ING_THREADS, 0, st>>>(raw_perm, n_raw);
        auto flags = thrust::make_transform_iterator(iota, HeadFlagCT{time, cam});
        VB_CHECK(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, flags, raw_pair, (int)n_raw, st));
    }
    VB_KERNEL_CHECK();
    int last = 0;
    VB_CHECK(read_back(&last, raw_pair + (n_raw - 1), sizeof(int), st));
    *h_n_pairs = (int64_t)last + 1;
    if (h_sorted) *h_sorted = unsorted ? 0 : 1;
    count_launches(unsorted ? 2 : 2);   // check_sorted + (make_keys | iota); CUB's sort / scan are library kernels
    return 0;
}

inline int64_t ingest_windows(int64_t n_edges, int64_t n_c, int64_t tile_len) {
    const int64_t per_cam = (n_edges + n_c - 1) / n_c;
    const int64_t w = (per_cam + tile_len - 1) / tile_len;
    return w < 1 ? 1 : w;
}

inline int64_t ingest_max_tiles(int64_t n_edges, int64_t n_c, int64_t tile_len) {
    return n_edges / tile_len + n_c * ingest_windows(n_edges, n_c, tile_len) + 2;
}

}  // namespace vb
