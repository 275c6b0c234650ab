// C ABI of the vican_b200 CUDA extension (see include/vican_b200.h).  sm_100a only.
#include "../../include/vican_b200.h"

#include "cg.cuh"
#include "common.cuh"
#include "evaluate.cuh"
#include "ingest.cuh"
#include "lobpcg.cuh"
#include "lsqr.cuh"
#include "nccl_shim.cuh"
#include "passes.cuh"
#include "peer.cuh"
#include "rotation.cuh"
#include "schur.cuh"
#include "translation.cuh"

using namespace vb;

namespace {

__global__ void se3_compose_kernel(const double* __restrict__ Ra, const double* __restrict__ ta, const double* __restrict__ Rb,
                                   const double* __restrict__ tb, double* __restrict__ Ro, double* __restrict__ to, int64_t n,
                                   int round_f32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double A[9], B[9], C[9], a[3], b[3], c[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) { A[k] = Ra[9 * i + k]; B[k] = Rb[9 * i + k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = ta[3 * i + k]; b[k] = tb[3 * i + k]; }
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

__global__ void se3_invert_kernel(const double* __restrict__ R, const double* __restrict__ t, double* __restrict__ Ri,
                                  double* __restrict__ ti, int64_t n, int round_f32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double A[9], a[3], c[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) A[k] = R[9 * i + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = t[3 * i + k];
    mtv3(A, a, c);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = A[3 * k + r];
            Ri[9 * i + 3 * r + k] = round_f32 ? (double)(float)v : v;
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double v = -c[k]; ti[3 * i + k] = round_f32 ? (double)(float)v : v; }
}

}  // namespace

extern "C" {

const char* vb_version(void) { return "vican_b200 0.1.0 (sm_100a)"; }

int vb_gather_stride(void) { return GSTRIDE; }

int64_t vb_launch_count(void) { return (int64_t)launch_counter().load(); }

const char* vb_status_string(int code) {
    if (code < 0) return cudaGetErrorString((cudaError_t)(-code));
    switch (code) {
        case VB_STATUS_OK: return "ok";
        case VB_STATUS_NOT_CONVERGED: return "conjugate gradient did not converge";
        case VB_STATUS_EIG_STALLED: return "eigen-iteration hit max_inner before reaching tol";
        case VB_STATUS_BAD_ARGUMENT: return "bad argument / workspace too small";
        case VB_STATUS_PEER_TIMEOUT: return "a peer rank never arrived at a peer-memory all-reduce (results invalid)";
        case VB_STATUS_SINGULAR: return "camera Schur complement is not positive definite (disconnected graph?)";
        default: return "unknown status";
    }
}

int vb_se3_compose_batch(const double* Ra, const double* ta, const double* Rb, const double* tb, double* Rout,
                         double* tout, int64_t n, int round_f32, void* stream) {
    if (n <= 0) return 0;
    se3_compose_kernel<<<node_grid(n), NODE_THREADS, 0, (cudaStream_t)stream>>>(Ra, ta, Rb, tb, Rout, tout, n, round_f32);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_se3_invert_batch(const double* R, const double* t, double* Rinv, double* tinv, int64_t n, int round_f32,
                        void* stream) {
    if (n <= 0) return 0;
    se3_invert_kernel<<<node_grid(n), NODE_THREADS, 0, (cudaStream_t)stream>>>(R, t, Rinv, tinv, n, round_f32);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_polar_so3_batch(const double* M, double* R, int64_t n, void* stream) {
    if (n <= 0) return 0;
    polar_batch_kernel<<<node_grid(n), NODE_THREADS, 0, (cudaStream_t)stream>>>(M, R, n);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_svd3_factors_batch(const double* M, double* rot, double* sym_pos, double* sym_inv, int64_t n, void* stream) {
    if (n <= 0) return 0;
    svd_factors_batch_kernel<<<node_grid(n), NODE_THREADS, 0, (cudaStream_t)stream>>>(M, rot, sym_pos, sym_inv, n);
    VB_KERNEL_CHECK();
    return 0;
}

// ----------------------------------------------------------------------------- evaluation
int64_t vb_gauge_workspace_bytes(int64_t n) { return (int64_t)ev_grid(n) * 12 * (int64_t)sizeof(double); }

int vb_optimize_gauge(const double* Ra, const double* ta, const double* Rb, const double* tb, int64_t n, double* gauge_R,
                      double* gauge_t, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n <= 0 || (ta == nullptr) != (tb == nullptr) || (ta != nullptr && gauge_t == nullptr)) return VB_STATUS_BAD_ARGUMENT;
    if (workspace_bytes < vb_gauge_workspace_bytes(n)) return VB_STATUS_BAD_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = ev_grid(n);
    gauge_partial_kernel<<<nb, EV_THREADS, 0, st>>>(Ra, ta, Rb, tb, n, (double*)workspace);
    gauge_finish_kernel<<<1, 32, 0, st>>>((const double*)workspace, nb, n, gauge_R, ta ? gauge_t : nullptr);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_distance_so3_batch(const double* R1, const double* R2, double* deg, int64_t n, void* stream) {
    if (n <= 0) return 0;
    distance_so3_kernel<<<(int)((n + EV_THREADS - 1) / EV_THREADS), EV_THREADS, 0, (cudaStream_t)stream>>>(R1, R2, deg, n);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_se3_left_compose_batch(const double* Rg, const double* tg, const double* R, const double* t, double* Rout,
                              double* tout, int64_t n, int round_f32, void* stream) {
    if (n <= 0) return 0;
    se3_left_compose_kernel<<<(int)((n + EV_THREADS - 1) / EV_THREADS), EV_THREADS, 0, (cudaStream_t)stream>>>(
        Rg, tg, R, t, Rout, tout, n, round_f32);
    VB_KERNEL_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------ ingestion
int64_t vb_ingest_workspace_bytes(int64_t n_raw) {
    IngestWork w = carve_ingest(nullptr, n_raw < 1 ? 1 : n_raw);
    return w.bytes;
}

int vb_ingest_sort(const int32_t* cam, const int32_t* time, int64_t n_raw, int64_t n_c, int64_t n_t, int32_t* raw_perm,
                   int32_t* raw_pair, int64_t* h_n_pairs, int32_t* h_sorted, void* workspace, int64_t workspace_bytes,
                   void* stream) {
    return ingest_sort(cam, time, n_raw, n_c, n_t, raw_perm, raw_pair, h_n_pairs, h_sorted, workspace, workspace_bytes,
                       (cudaStream_t)stream);
}

int64_t vb_ingest_max_tiles(int64_t n_edges, int64_t n_c, int64_t tile_len) {
    return ingest_max_tiles(n_edges, n_c, tile_len);
}

int64_t vb_ingest_windows(int64_t n_edges, int64_t n_c, int64_t tile_len) { return ingest_windows(n_edges, n_c, tile_len); }

int vb_ingest_build(const int32_t* cam, const int32_t* time, const int32_t* marker, const double* R, const double* k_r,
                    const double* k_t, const double* markerC, int64_t n_raw, int round_kr_f32, const int32_t* raw_perm,
                    const int32_t* raw_pair, int64_t n_pairs, int64_t n_c, int64_t n_t, int64_t tile_len,
                    int32_t* t_rowptr, int32_t* t_cam, int32_t* t_time, double* t_B, double* t_a, double* t_w,
                    int32_t* pair_start, int32_t* c_segptr, int32_t* c_time, double* c_B, double* c_w,
                    int32_t* c_order, int32_t* tile_cam, int32_t* tile_start, int32_t* tile_off, int64_t* h_n_tiles, double* deg_t,
                    double* deg_c, const vb_arrival* arrival, int64_t n_markers, int identity_perm, void* workspace,
                    int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n_raw <= 0 || n_pairs <= 0 || tile_len <= 0) return VB_STATUS_BAD_ARGUMENT;
    if (arrival != nullptr && (arrival->n_chunks < 1 || arrival->n_chunks > 64)) return VB_STATUS_BAD_ARGUMENT;
    const int64_t E = n_pairs;
    const int64_t n_seg_ws = ingest_windows(E, n_c, tile_len) * n_c + 1;   // a chunk can hold fewer detections than cameras
    IngestWork w = carve_ingest(workspace, n_raw > n_seg_ws ? n_raw : n_seg_ws);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    pair_start_kernel<<<ing_grid(n_raw), ING_THREADS, 0, st>>>(raw_pair, pair_start, n_raw, E);
    // camera-pass order first ((time window, camera, time); tiles = runs of (window, camera)), so that the fold can
    // write both copies of the blocks in one pass over the detections
    const int64_t n_win = ingest_windows(E, n_c, tile_len);
    const int64_t n_seg = n_win * n_c;
    uint32_t* keys32_a = (uint32_t*)w.keys_a;
    uint32_t* keys32_b = (uint32_t*)w.keys_b;
    int* c_pos = w.tmp_a;
    pair_keys_kernel<<<ing_grid(E), ING_THREADS, 0, st>>>(cam, time, raw_perm, pair_start, E, n_c, n_t, n_win, t_cam, t_time,
                                                         keys32_a, w.vals_a);
    size_t tb = w.cub_bytes;
    VB_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, (const uint32_t*)keys32_a, keys32_b, (const int*)w.vals_a,
                                             c_order, (int)E, 0, key_bits(n_c, n_win), st));
    seg_ptr_kernel<<<ing_grid(E), ING_THREADS, 0, st>>>((const int*)keys32_b, nullptr, c_segptr, E, n_seg);
    inverse_perm_kernel<<<ing_grid(E), ING_THREADS, 0, st>>>(c_order, c_pos, E);
    // The fold is the only step that reads the rotations (72 of the 92 bytes per detection).  When they are still
    // crossing PCIe (vb_arrival: chunks of consecutive detections, one event each) and the detections arrived sorted
    // (sorted position = raw index), every chunk is folded as soon as it has landed, under the copy of the next.
    int64_t h_pair_end[64];
    int n_fold = 1;
    h_pair_end[0] = E;
    if (arrival != nullptr && arrival->sorted_input && arrival->n_chunks > 1) {
        n_fold = arrival->n_chunks;
        int64_t* d_tmp = (int64_t*)w.keys_a;   // keys_a is free again (the sort has consumed it)
        VB_CHECK(cudaMemcpyAsync(d_tmp, arrival->h_raw_end, n_fold * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        chunk_pairs_kernel<<<1, 64, 0, st>>>(raw_pair, d_tmp, n_fold, n_raw, E, d_tmp + 64);
        VB_CHECK(read_back(h_pair_end, d_tmp + 64, n_fold * sizeof(int64_t), st));
        h_pair_end[n_fold - 1] = E;
    }
    int64_t p_lo = 0;
    for (int kf = 0; kf < n_fold; ++kf) {
        if (arrival != nullptr) {
            if (n_fold > 1) VB_CHECK(cudaStreamWaitEvent(st, (cudaEvent_t)arrival->h_events[kf], 0));
            else for (int q = 0; q < arrival->n_chunks; ++q) VB_CHECK(cudaStreamWaitEvent(st, (cudaEvent_t)arrival->h_events[q], 0));
        }
        const int64_t p_hi = h_pair_end[kf];
        if (p_hi > p_lo)
            fold_both_kernel<<<(int)((p_hi - p_lo + FOLD_THREADS - 1) / FOLD_THREADS), FOLD_THREADS, 0, st>>>(
                marker, R, k_r, k_t, markerC, round_kr_f32, raw_perm, pair_start, p_lo, p_hi, t_time, c_pos, t_B, t_a, t_w, c_B, c_time, c_w,
                (int)(n_markers > 0x7fffffff ? 0x7fffffff : n_markers), identity_perm);
        p_lo = p_hi > p_lo ? p_hi : p_lo;
    }
    seg_ptr_kernel<<<ing_grid(E), ING_THREADS, 0, st>>>(t_time, nullptr, t_rowptr, E, n_t);
    seg_sum_kernel<<<ing_grid(n_t * 32), ING_THREADS, 0, st>>>(t_rowptr, nullptr, t_a, deg_t, n_t);
    cam_runs_sum_kernel<<<ing_grid(n_c * 32), ING_THREADS, 0, st>>>(c_segptr, n_win, n_c, c_order, t_a, deg_c);
    VB_KERNEL_CHECK();
    tile_count_kernel<<<ing_grid(n_seg + 1), ING_THREADS, 0, st>>>(c_segptr, w.tmp_c, n_seg, (int)tile_len);
    tb = w.cub_bytes;
    VB_CHECK(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, (const int*)w.tmp_c, tile_off, (int)(n_seg + 1), st));
    tile_fill_kernel<<<ing_grid(n_seg), ING_THREADS, 0, st>>>(c_segptr, tile_off, tile_cam, tile_start, n_seg, n_c, (int)tile_len);
    VB_KERNEL_CHECK();
    int last_off = 0;
    VB_CHECK(read_back(&last_off, tile_off + n_seg, sizeof(int), st));
    const int64_t n_tiles = (int64_t)last_off;
    *h_n_tiles = n_tiles;
    count_launches(9 + n_fold);   // pair_start, pair_keys, seg_ptr x2, inverse_perm, fold_both (per chunk), seg_sum, cam_runs_sum, tile_count, tile_fill
    return 0;
}

int vb_count_components(const vb_graph* g, const int32_t* t_time, int32_t* labels, int64_t* h_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = g->n_c + g->n_t, E = g->n_edges;
    if (n <= 0 || E <= 0) return VB_STATUS_BAD_ARGUMENT;
    int* flag = labels + n;   // labels holds n + 2 ints: labels, changed flag, component count
    iota_kernel<<<ing_grid(n), ING_THREADS, 0, st>>>(labels, n);
    int launches = 1;
    for (int round = 0; round < 10000; ++round) {
        VB_CHECK(cudaMemsetAsync(flag, 0, 2 * sizeof(int), st));
        cc_hook_kernel<<<ing_grid(E), ING_THREADS, 0, st>>>(g->t_cam, t_time, E, (int)g->n_c, labels, flag);
        cc_jump_kernel<<<ing_grid(n), ING_THREADS, 0, st>>>(labels, n);
        launches += 2;
        int changed = 0;
        VB_CHECK(read_back(&changed, flag, sizeof(int), st));
        if (!changed) break;
    }
    cc_count_kernel<<<ing_grid(n), ING_THREADS, 0, st>>>(labels, n, flag + 1);
    int count = 0;
    VB_CHECK(read_back(&count, flag + 1, sizeof(int), st));
    VB_KERNEL_CHECK();
    count_launches(launches + 1);
    *h_count = count;
    return 0;
}

// dst[i] = src[i] + add: appends an index array of a freshly ingested chunk behind the arrays of a growing graph
__global__ void offset_copy_kernel(int* __restrict__ dst, const int* __restrict__ src, int64_t n, int add) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}
__global__ void add_inplace_kernel(double* __restrict__ dst, const double* __restrict__ src, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

int vb_offset_copy_i32(int32_t* dst, const int32_t* src, int64_t n, int32_t add, void* stream) {
    if (n <= 0) return 0;
    offset_copy_kernel<<<ing_grid(n), ING_THREADS, 0, (cudaStream_t)stream>>>(dst, src, n, add);
    VB_KERNEL_CHECK();
    count_launches(1);
    return 0;
}

int vb_add_inplace_f64(double* dst, const double* src, int64_t n, void* stream) {
    if (n <= 0) return 0;
    add_inplace_kernel<<<ing_grid(n), ING_THREADS, 0, (cudaStream_t)stream>>>(dst, src, n);
    VB_KERNEL_CHECK();
    count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------- rotation
int vb_pass_time(const vb_graph* g, int mode, const double* X, const double* lamT, double* out, void* stream) {
    return launch_pass_time(mode, g->t_rowptr, g->t_cam, g->t_B, X, lamT, out, g->n_t, (cudaStream_t)stream);
}

int vb_pass_cam(const vb_graph* g, const double* W, double* Y, void* stream) {
    return launch_pass_cam(g, W, Y, (cudaStream_t)stream);
}

int vb_pad_blocks(const double* src9, double* dst12, int64_t n, void* stream) {
    return launch_pad_blocks(src9, dst12, n, (cudaStream_t)stream);
}

int vb_primal_update(const double* M, double* r_c, double* lamC, double* lamCinv, int64_t n_c, void* stream) {
    if (n_c <= 0) return 0;
    primal_update_kernel<<<node_grid(n_c), NODE_THREADS, 0, (cudaStream_t)stream>>>(M, r_c, lamC, lamCinv, n_c);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_dual_update(const double* Yt, double* r_t, double* lamT, double* Wt, int64_t n_t, void* stream) {
    if (n_t <= 0) return 0;
    dual_update_kernel<<<node_grid(n_t), NODE_THREADS, 0, (cudaStream_t)stream>>>(Yt, r_t, lamT, Wt, n_t);
    VB_KERNEL_CHECK();
    return 0;
}

int vb_gauge_project(const double* V, double* r_c, int64_t n_c, void* stream) {
    if (n_c <= 0) return 0;
    gauge_project_kernel<<<node_grid(n_c), NODE_THREADS, 0, (cudaStream_t)stream>>>(V, r_c, n_c);
    VB_KERNEL_CHECK();
    return 0;
}

int64_t vb_so3sync_workspace_bytes(int64_t n_c, int64_t n_t) { return carve_so3(nullptr, n_c, n_t).bytes; }

int vb_so3sync_run(const vb_graph* g, const vb_so3_options* opt, double* r_c, double* r_t, void* workspace,
                   int64_t workspace_bytes, vb_so3_stats* stats, void* stream) {
    vb_so3_stats local;
    if (stats == nullptr) stats = &local;
    const int rc = so3sync_run(g, opt, r_c, r_t, workspace, workspace_bytes, stats, (cudaStream_t)stream);
    count_launches(stats->kernel_launches);
    return rc;
}

// ---------------------------------------------------------------------------- translation
int vb_trans_rhs(const vb_graph* g, const int32_t* raw_perm, const int32_t* pair_start, const int32_t* marker,
                 const double* t_cm, const double* k_t, const double* marker_q, const double* r_c, const double* r_t,
                 const int32_t* t_time, double* pair_g, double* d_sorted, double* rhs_c, double* rhs_t, double* r_c_pad,
                 int64_t n_markers, int identity_perm, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t E = g->n_edges;
    if (E <= 0 || r_c_pad == nullptr) return VB_STATUS_BAD_ARGUMENT;
    { int rc = launch_pad_blocks(r_c, r_c_pad, g->n_c, st); if (rc) return rc; }
    trans_pair_kernel<<<tr_grid(E), TR_THREADS, 0, st>>>(raw_perm, pair_start, marker, t_cm, k_t, marker_q, r_c_pad, r_t,
                                                        g->t_cam, t_time, E, pair_g, d_sorted, (int)(n_markers > 0x7fffffff ? 0x7fffffff : n_markers),
                                                        identity_perm);
    VB_CHECK(cudaMemsetAsync(rhs_c, 0, 3 * g->n_c * sizeof(double), st));
    VB_CHECK(cudaMemsetAsync(rhs_t, 0, 3 * g->n_t * sizeof(double), st));
    seg_sum3_kernel<<<tr_warp_grid(g->n_t), TR_THREADS, 0, st>>>(g->t_rowptr, nullptr, pair_g, 1.0, rhs_t, g->n_t);
    cam_runs_sum3_kernel<<<tr_warp_grid(g->n_c), TR_THREADS, 0, st>>>(g->c_segptr, g->n_windows, g->n_c, g->c_order, pair_g, -1.0, rhs_c);
    VB_KERNEL_CHECK();
    count_launches(4);
    return 0;
}

int64_t vb_trans_cg_workspace_bytes(int64_t n_c, int64_t n_t) { return carve_cg(nullptr, n_c, n_t).bytes; }

int64_t vb_sell_workspace_bytes(int64_t n_c, int64_t n_t, int64_t n_windows) { return carve_sell(nullptr, n_c, n_t, n_windows).bytes; }

int vb_sell_count(const vb_graph* g, int32_t* st_ptr, int32_t* sc_ptr, int64_t* h_chunks_t, int64_t* h_chunks_c,
                  void* workspace, int64_t workspace_bytes, void* stream) {
    return sell_count(g, st_ptr, sc_ptr, h_chunks_t, h_chunks_c, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vb_sell_fill(const vb_graph* g, const int32_t* st_ptr, int32_t* st_idx, double* st_w, const int32_t* sc_ptr,
                 int32_t* sc_idx, double* sc_w, int64_t chunks_c, void* workspace, int64_t workspace_bytes, void* stream) {
    return sell_fill(g, st_ptr, st_idx, st_w, sc_ptr, sc_idx, sc_w, chunks_c, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vb_trans_cg(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t, double rtol,
                int64_t maxiter, int jacobi, const int32_t* unk_c, const int32_t* unk_t, int32_t* h_iters, void* workspace,
                int64_t workspace_bytes, vb_allreduce_fn allreduce, void* allreduce_ctx, int owns_camera_diagonal,
                void* stream) {
    return trans_cg(g, rhs_c, rhs_t, x_c, x_t, rtol, maxiter, jacobi, unk_c, unk_t, h_iters, workspace, workspace_bytes,
                    allreduce, allreduce_ctx, owns_camera_diagonal, (cudaStream_t)stream);
}

int64_t vb_trans_lsqr_workspace_bytes(int64_t n_c, int64_t n_t, int64_t n_raw) {
    return carve_lsqr(nullptr, n_c, n_t, n_raw).bytes;
}

int vb_trans_lsqr(const vb_graph* g, const int32_t* raw_perm, const int32_t* raw_pair, const int32_t* pair_start,
                  const int32_t* t_time, const double* k_t, const double* d_sorted, int64_t n_raw, double* x_c,
                  double* x_t, double atol, double btol, double conlim, int64_t iter_lim, int32_t* h_istop,
                  int32_t* h_iters, void* workspace, int64_t workspace_bytes, vb_allreduce_fn allreduce,
                  void* allreduce_ctx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_c = g->n_c, n_t = g->n_t;
    LsqrWork w = carve_lsqr(workspace, n_c, n_t, n_raw);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    double* hs = pinned_status();
    const bool multi = allreduce != nullptr;
    // edge-sharded runs: rows (detections) and time nodes are local, the camera block is replicated.  The three
    // local norms are collapsed to scalars red[0..2] = ||u||^2, ||v_t||^2, ||w_t||^2 and summed over the ranks;
    // the camera part of A^T u is summed as a vector.  3 collectives per bidiagonalisation step.
    double* red = w.partial + 2048;
    lsqr_clear_kernel<<<1, 1, 0, st>>>(w.sc);
    lsqr_rows_kernel<<<tr_grid(n_raw), TR_THREADS, 0, st>>>(raw_perm, raw_pair, g->t_cam, t_time, k_t, d_sorted, n_raw,
                                                           w.row_cam, w.row_time, w.kt_sorted, w.u, w.keys_a, w.vals_a);
    VB_KERNEL_CHECK();
    size_t tb = w.cub_bytes;
    int cbits = 1;
    while (cbits < 64 && ((uint64_t)n_c >> cbits) != 0) ++cbits;
    VB_CHECK(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, (const uint64_t*)w.keys_a, w.keys_b, (const int*)w.vals_a,
                                             w.cam_rows, (int)n_raw, 0, cbits, st));
    seg_ptr_kernel<<<ing_grid(n_raw), ING_THREADS, 0, st>>>(w.row_cam, w.cam_rows, w.cam_ptr, n_raw, n_c);
    VB_KERNEL_CHECK();
    const int nb_u = sumsq_grid(3 * n_raw), nb_c = sumsq_grid(3 * n_c), nb_t = sumsq_grid(3 * n_t);
    long long launches = 3;
    auto u_norm = [&](int init) -> int {   // beta = ||u||
        sumsq_partial_kernel<<<nb_u, TR_THREADS, 0, st>>>(w.u, 3 * n_raw, 1.0, w.partial);
        if (multi) {
            lsqr_collapse_kernel<<<1, 1, 0, st>>>(w.partial, nb_u, red);
            int rc = allreduce(allreduce_ctx, red, 3, (void*)st);
            if (rc) return rc;
            lsqr_set_kernel<<<1, 1, 0, st>>>(w.partial, red);
            launches += 2;
        }
        lsqr_s1_kernel<<<1, 1, 0, st>>>(w.sc, w.partial, multi ? 1 : nb_u, init);
        launches += 2;
        return 0;
    };
    auto v_step = [&](int init, bool with_w) -> int {   // v = A^T u - beta v, alfa = ||v||
        lsqr_vt_kernel<<<tr_warp_grid(n_t), TR_THREADS, 0, st>>>(g->t_rowptr, pair_start, w.kt_sorted, w.u, w.v_t, n_t, w.sc, init);
        lsqr_vc_kernel<<<tr_warp_grid(n_c), TR_THREADS, 0, st>>>(w.cam_ptr, w.cam_rows, w.kt_sorted, w.u, w.v_c, n_c, w.sc, init,
                                                                 multi ? w.acc_c : nullptr);
        launches += 2;
        if (multi) {
            int rc = allreduce(allreduce_ctx, w.acc_c, 3 * n_c, (void*)st);
            if (rc) return rc;
            lsqr_vc_finish_kernel<<<tr_grid(3 * n_c), TR_THREADS, 0, st>>>(w.acc_c, w.v_c, 3 * n_c, w.sc, init);
            // local time parts of ||v||^2 and (when the step needs it) ||w_old||^2, one collective
            sumsq_partial_kernel<<<nb_t, TR_THREADS, 0, st>>>(w.v_t, 3 * n_t, 1.0, w.partial + 1024);
            lsqr_collapse_kernel<<<1, 1, 0, st>>>(w.partial + 1024, nb_t, red + 1);
            if (with_w) {
                sumsq_partial_kernel<<<nb_t, TR_THREADS, 0, st>>>(w.w_t, 3 * n_t, 1.0, w.partial + 1024);
                lsqr_collapse_kernel<<<1, 1, 0, st>>>(w.partial + 1024, nb_t, red + 2);
            }
            rc = allreduce(allreduce_ctx, red, 3, (void*)st);   // red[0] is summed again: it is not read any more
            if (rc) return rc;
            lsqr_set_kernel<<<1, 1, 0, st>>>(w.partial + 1024, red + 1);
            launches += with_w ? 6 : 4;
        } else {
            sumsq_partial_kernel<<<nb_t, TR_THREADS, 0, st>>>(w.v_t, 3 * n_t, 1.0, w.partial + 1024);
            launches += 1;
        }
        sumsq_partial_kernel<<<nb_c, TR_THREADS, 0, st>>>(w.v_c, 3 * n_c, 1.0, w.partial);
        lsqr_s2_kernel<<<1, 1, 0, st>>>(w.sc, w.partial, nb_c, multi ? 1 : nb_t, init);
        launches += 2;
        VB_KERNEL_CHECK();
        return 0;
    };
    // initial bidiagonalisation vectors (lsqr.py:372-398)
    { int rc = u_norm(1); if (rc) return rc; }
    { int rc = v_step(1, false); if (rc) return rc; }
    lsqr_x_kernel<<<tr_grid(3 * n_c), TR_THREADS, 0, st>>>(w.v_c, w.w_c, x_c, 3 * n_c, w.sc, 1);
    lsqr_x_kernel<<<tr_grid(3 * n_t), TR_THREADS, 0, st>>>(w.v_t, w.w_t, x_t, 3 * n_t, w.sc, 1);
    VB_KERNEL_CHECK();
    launches += 2;
    for (int64_t it = 0; it < iter_lim; ++it) {
        lsqr_u_kernel<<<tr_grid(n_raw), TR_THREADS, 0, st>>>(w.row_cam, w.row_time, w.kt_sorted, w.v_c, w.v_t, w.u, n_raw, w.sc);
        { int rc = u_norm(0); if (rc) return rc; }
        { int rc = v_step(0, true); if (rc) return rc; }
        sumsq_partial_kernel<<<nb_c, TR_THREADS, 0, st>>>(w.w_c, 3 * n_c, 1.0, w.partial);
        if (multi) lsqr_set_kernel<<<1, 1, 0, st>>>(w.partial + 1024, red + 2);
        else sumsq_partial_kernel<<<nb_t, TR_THREADS, 0, st>>>(w.w_t, 3 * n_t, 1.0, w.partial + 1024);
        lsqr_x_kernel<<<tr_grid(3 * n_c), TR_THREADS, 0, st>>>(w.v_c, w.w_c, x_c, 3 * n_c, w.sc, 0);
        lsqr_x_kernel<<<tr_grid(3 * n_t), TR_THREADS, 0, st>>>(w.v_t, w.w_t, x_t, 3 * n_t, w.sc, 0);
        lsqr_s3_kernel<<<1, 1, 0, st>>>(w.sc, w.partial, nb_c, multi ? 1 : nb_t, atol, btol, conlim, (double)iter_lim);
        VB_KERNEL_CHECK();
        launches += 6;
        VB_CHECK(cudaMemcpyAsync(hs, w.sc, LS_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
        VB_CHECK(cudaStreamSynchronize(st));
        if (hs[LS_ISTOP] != 0.0) break;
    }
    VB_CHECK(cudaMemcpyAsync(hs, w.sc, LS_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
    VB_CHECK(cudaStreamSynchronize(st));
    if (h_istop) *h_istop = (int32_t)hs[LS_ISTOP];
    if (h_iters) *h_iters = (int32_t)hs[LS_ITN];
    count_launches(launches);
    return 0;
}

int64_t vb_trans_schur_workspace_bytes(int64_t n_c, int64_t n_t) { return carve_schur(nullptr, n_c, n_t).bytes; }

int vb_trans_schur_direct(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t,
                          void* workspace, int64_t workspace_bytes, void* stream) {
    return trans_schur_direct(g, rhs_c, rhs_t, x_c, x_t, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------ multi-GPU
int vb_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int vb_nccl_unique_id(void* h_out, int64_t bytes) {
    if (!nccl_api().ok || bytes < (int64_t)sizeof(ncclUniqueId)) return VB_STATUS_BAD_ARGUMENT;
    ncclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) return VB_STATUS_BAD_ARGUMENT;
    memcpy(h_out, &id, sizeof(id));
    return 0;
}

int vb_nccl_init(const void* h_id, int64_t bytes, int rank, int nranks, void** ctx_out) {
    if (!nccl_api().ok || bytes < (int64_t)sizeof(ncclUniqueId)) return VB_STATUS_BAD_ARGUMENT;
    ncclUniqueId id;
    memcpy(&id, h_id, sizeof(id));
    NcclCtx* ctx = new NcclCtx{nullptr, rank, nranks};
    if (nccl_api().CommInitRank(&ctx->comm, nranks, id, rank) != ncclSuccess) { delete ctx; return VB_STATUS_BAD_ARGUMENT; }
    *ctx_out = ctx;
    return 0;
}

int vb_nccl_destroy(void* ctx) {
    if (!ctx) return 0;
    NcclCtx* c = (NcclCtx*)ctx;
    nccl_api().CommDestroy(c->comm);
    delete c;
    return 0;
}

int vb_nccl_allreduce(void* ctx, double* buf, int64_t count, void* stream) {
    NcclCtx* c = (NcclCtx*)ctx;
    ncclResult_t r = nccl_api().AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, c->comm, (cudaStream_t)stream);
    return r == ncclSuccess ? 0 : VB_STATUS_BAD_ARGUMENT;
}

void* vb_nccl_allreduce_fn(void) { return (void*)&vb_nccl_allreduce; }

// ---- peer-memory windows (CUDA IPC over NVLink) ----
int vb_peer_create(int rank, int nranks, int64_t capacity_doubles, void** ctx_out, void* h_handle_out64) {
    if (nranks < 1 || nranks > PEER_MAX || rank < 0 || rank >= nranks || capacity_doubles < 2) return VB_STATUS_BAD_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    static_assert(sizeof(PeerCtrl) <= PEER_CTRL_BYTES, "control block");
    PeerCtx* ctx = new PeerCtx();
    const int64_t cap = (capacity_doubles + 31) & ~(int64_t)31;   // even (16-byte loads) and 256-byte aligned buffers
    ctx->bytes = PEER_CTRL_BYTES + 2 * (size_t)cap * sizeof(double);
    cudaError_t e = cudaMalloc(&ctx->base, ctx->bytes);
    if (e == cudaSuccess) e = cudaMemset(ctx->base, 0, ctx->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ctx->base);
    if (e != cudaSuccess) { if (ctx->base) cudaFree(ctx->base); delete ctx; return -(int)e; }
    memcpy(h_handle_out64, &h, sizeof(h));
    ctx->dev.rank = rank; ctx->dev.world = nranks; ctx->dev.cap = cap;
    for (int r = 0; r < PEER_MAX; ++r) { ctx->dev.ctrl[r] = nullptr; ctx->dev.buf[r] = nullptr; }
    ctx->dev.ctrl[rank] = (PeerCtrl*)ctx->base;
    ctx->dev.buf[rank] = (double*)((char*)ctx->base + PEER_CTRL_BYTES);
    *ctx_out = ctx;
    return 0;
}

int vb_peer_connect(void* vctx, const void* h_all_handles) {
    PeerCtx* ctx = (PeerCtx*)vctx;
    if (!ctx) return VB_STATUS_BAD_ARGUMENT;
    for (int r = 0; r < ctx->dev.world; ++r) {
        if (r == ctx->dev.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)h_all_handles + 64 * (size_t)r, sizeof(h));
        void* p = nullptr;
        VB_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->opened[r] = p;
        ctx->dev.ctrl[r] = (PeerCtrl*)p;
        ctx->dev.buf[r] = (double*)((char*)p + PEER_CTRL_BYTES);
    }
    return 0;
}

int vb_peer_destroy(void* vctx) {
    PeerCtx* ctx = (PeerCtx*)vctx;
    if (!ctx) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < PEER_MAX; ++r)
        if (ctx->opened[r]) cudaIpcCloseMemHandle(ctx->opened[r]);
    if (ctx->base) cudaFree(ctx->base);
    delete ctx;
    return 0;
}

int vb_peer_allreduce(void* ctx, double* buf, int64_t count, void* stream) {
    return launch_peer_allreduce((PeerCtx*)ctx, buf, count, (cudaStream_t)stream);
}

void* vb_peer_allreduce_fn(void) { return (void*)&vb_peer_allreduce; }

int vb_peer_stamps(void* vctx, uint64_t* h_out4, void* stream) {
    PeerCtx* ctx = (PeerCtx*)vctx;
    if (!ctx) return VB_STATUS_BAD_ARGUMENT;
    VB_CHECK(cudaMemcpyAsync(h_out4, ((PeerCtrl*)ctx->base)->stamp, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VB_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int vb_peer_status(void* vctx, void* stream) {
    PeerCtx* ctx = (PeerCtx*)vctx;
    if (!ctx) return VB_STATUS_BAD_ARGUMENT;
    unsigned int t = 0;
    VB_CHECK(cudaMemcpyAsync(&t, &((PeerCtrl*)ctx->base)->timeouts, sizeof(t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VB_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return t == 0 ? VB_STATUS_OK : VB_STATUS_PEER_TIMEOUT;
}

}  // extern "C"
