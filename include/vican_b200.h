/* vican_b200 -- C ABI of the B200-native bipartite SE(3) synchronisation solver.
 *
 * The reference (gabmoreira/vican) is pure Python and has NO plugin / FFI boundary for this
 * path: its boundary is the two Python functions
 *     vican/bipgo.py:353  bipartite_se3sync(...)
 *     vican/bipgo.py:493  object_bipartite_se3sync(...)
 * The drop-in keeps those signatures in Python (vican_b200/bipgo.py) and calls the entry
 * points below through ctypes.  Each entry point names the reference lines it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors) unless the name
 *    starts with h_; the library never allocates device memory (workspace sizes are queried);
 *  - `stream` is a cudaStream_t passed as void*; work is asynchronous on it unless stated;
 *  - return value: 0 = OK, < 0 = -(cudaError_t), > 0 = solver status (VB_STATUS_*);
 *  - 3x3 blocks are row-major, 9 doubles; all arithmetic is fp64.
 */
#ifndef VICAN_B200_H
#define VICAN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_STATUS_OK 0
#define VB_STATUS_NOT_CONVERGED 1   /* CG hit maxiter (reference: assert exit_code == 0, bipgo.py:478) */
#define VB_STATUS_EIG_STALLED 2     /* LOBPCG hit max_inner before tol (result still returned) */
#define VB_STATUS_BAD_ARGUMENT 3
#define VB_STATUS_PEER_TIMEOUT 5     /* peer-memory all-reduce: a rank never arrived (results invalid) */
#define VB_STATUS_SINGULAR 4         /* dense direct solve: Schur complement not positive definite (disconnected graph) */

/* Device-resident bipartite graph of aggregated (camera, time) edges, stored twice:
 * sorted by time node (CSR) and sorted by camera (CSC) with camera tiles for the camera
 * pass.  Built by vb_ingest_* from the raw detections (replaces the dict / COO / CSR
 * assembly of bipgo.py:203-276). */
typedef struct vb_graph {
    int64_t n_c;            /* cameras (global) */
    int64_t n_t;            /* time nodes (local to this rank) */
    int64_t n_edges;        /* aggregated edges E (local) */
    int64_t n_tiles;        /* camera tiles */
    int64_t n_windows;      /* time windows of the camera-pass order */
    const int32_t* t_rowptr;   /* [n_t+1] */
    const int32_t* t_cam;      /* [E]   camera of each time-sorted edge */
    const double*  t_B;        /* [E][9] block  sum k_r R_cm R_m^T R_0 */
    const double*  t_w;        /* [E]   sum k_t^2   (translation Laplacian weight) */
    const int32_t* c_segptr;   /* [n_windows*n_c+1] runs of the camera-pass order: run (w, c) = [c_segptr[w*n_c+c], +1) */
    const int32_t* c_order;    /* [E]   camera-pass position -> time-sorted edge */
    const int32_t* c_time;     /* [E]   time node of each camera-pass edge ((window, camera, time) order) */
    const double*  c_B;        /* [E][9] the same blocks TRANSPOSED (B_e^T), camera-pass order */
    const double*  c_w;        /* [E] */
    const int32_t* tile_cam;   /* [n_tiles] */
    const int32_t* tile_start; /* [n_tiles] */
    const int32_t* tile_off;   /* [n_windows*n_c+1] first tile of every (window, camera) run */
    double*        tile_part;  /* [n_tiles][9] scratch of the camera pass: per-tile sums, combined per camera in a
                                * fixed order (no atomics: the pass is bitwise reproducible) */
    const double*  deg_t;      /* [n_t]  sum of k_r over the node's edges (bipgo.py:271) */
    const double*  deg_c;      /* [n_c]  sum of k_r over the camera's LOCAL edges (bipgo.py:275) */
    /* Translation Laplacian J^T J (bipgo.py:471-477) in a sliced-ELL layout, both sides: slices of 8 rows,
     * chunks of 4 columns = 32 slots (slot 4 j + s of chunk q = column 4 q + s of row j of the slice; padding:
     * index -1, weight 0).  Built by vb_sell_count / vb_sell_fill; NULL until then (only vb_trans_cg needs it). */
    const int32_t* st_ptr;     /* [ceil(n_t/8)+1] first chunk of each slice of time rows */
    const int32_t* st_idx;     /* [32 * chunks_t] camera of the slot (ascending within a row) */
    const double*  st_w;       /* [32 * chunks_t] sum k_t^2 of the pair */
    const int32_t* sc_ptr;     /* [ceil(n_c/8)+1] camera rows */
    const int32_t* sc_idx;     /* [32 * chunks_c] time node of the slot (ascending within a row) */
    const double*  sc_w;       /* [32 * chunks_c] */
} vb_graph;

/* Optional cross-rank reduction hook (edge-sharded multi-GPU): called on `stream` after every
 * camera pass with the camera-side accumulator.  NULL on a single GPU. */
typedef int (*vb_allreduce_fn)(void* ctx, double* buf, int64_t count, void* stream);

typedef struct vb_so3_options {
    int32_t maxiter;        /* primal-dual iterations (bipgo.py:282) */
    int32_t max_inner;      /* cap on LOBPCG steps per outer iteration */
    double  tol;            /* eigen-residual tolerance relative to rms(Lambda_C) */
    vb_allreduce_fn allreduce;
    void*   allreduce_ctx;
    int32_t profile_events; /* != 0: bracket every edge-pass launch with CUDA events (stats->*_pass_ms) */
    int32_t no_shortcut;    /* != 0: always run the primal multiply as two edge passes (see stats->shortcut_outer) */
    int32_t identity_start; /* != 0: first eigen-solve starts from identity blocks instead of the one-hop spanning
                             * estimate project_SO3((P Lambda_T P^T E_0)_c) (2 extra edge passes, ~half the steps) */
    int32_t eval_gap;       /* != 0: also compute the next three eigenvalues (lambda_4..6) in every outer iteration, by a
                             * second eigen-solve on the operator deflated by the converged eigenvectors -- the
                             * reference's diagnostics (evals, eigengap = |lambda_4 / lambda_3|, bipgo.py:291, :336-339)
                             * and its early exit `max |lambda_1..5| <= 1e-6` (bipgo.py:283, :292), which can only
                             * fire on (nearly) disconnected graphs.  Costs one more eigen-solve per iteration. */
    double  tol_early;      /* > 0: eigen-residual tolerance of the outer iterations that are followed by at least
                             * `early_margin` more: inexact inner solves while the outer iteration is far from its
                             * fixed point.  The last `early_margin` iterations always use `tol`; if the last two of
                             * them do not accept their start block at the first step (the outer iteration had not
                             * reached its fixed point to `tol` two iterations before the end, so the history still
                             * matters), stats->inexact_unverified is set and the caller must repeat the run with
                             * tol_early = 0.  0: `tol` everywhere. */
    int32_t early_margin;
    int32_t reserved;
    void*   peer_ctx;       /* vb_peer_create context: the camera pass runs FUSED with its cross-rank sum over
                             * NVLink peer memory (allreduce / allreduce_ctx are then used for the few other
                             * reductions only and may point at vb_peer_allreduce); NULL: separate collective */
} vb_so3_options;

typedef struct vb_so3_stats {
    int32_t outer_done;
    int32_t time_passes;    /* launches of the time-sorted edge pass */
    int32_t cam_passes;     /* launches of the camera-sorted edge pass */
    int32_t lobpcg_steps;   /* cooperative LOBPCG kernels */
    int32_t kernel_launches;
    int32_t stalled_outer;  /* outer iterations whose eigen-solve hit max_inner */
    double  theta[3];       /* last Ritz values (the reference's evals0..2, bipgo.py:336-339) */
    double  resid[3];       /* last eigen-residual norms */
    double  anorm;
    int32_t inner_per_outer[64];
    /* filled when opt->profile_events: device time of the EXECUTED edge passes (speculative launches
     * that were skipped on the device are excluded), measured with CUDA events on the solver's stream */
    double  time_pass_ms;   /* sum over executed time passes (modes 0 and 1) */
    double  cam_pass_ms;    /* sum over executed camera passes */
    int32_t time_pass_timed;
    int32_t cam_pass_timed;
    /* outer iterations whose primal multiply was obtained WITHOUT edge passes: when the eigen-iteration accepts
     * its start block R (the previous r_c) at the first step, project_SO3(V_c V_0^-1) = R_c R_0^T exactly and
     * P Lambda_T P^T r_c = Y R_0^T with Y already computed for the eigen-residual (bipgo.py:295-300) */
    int32_t shortcut_outer;
    int32_t early_exit;     /* 1: the loop stopped before maxiter because max |lambda_1..5| <= 1e-6 (eval_gap only) */
    /* eval_gap: the five eigenvalues nearest zero of every outer iteration (bipgo.py:288-292), first 64 iterations */
    double  evals_hist[64][5];
    int32_t inexact_unverified;   /* see vb_so3_options.tol_early */
    int32_t reserved3;
} vb_so3_stats;

const char* vb_version(void);
/* Number of this library's own kernels launched (and executed) so far in this process: a tally kept by the
 * host wrappers (CUB's sort / scan kernels and speculative launches that return at once are not counted). */
int64_t vb_launch_count(void);
const char* vb_status_string(int code);

/* ---- batched geometry (vican/geometry.py) ------------------------------------------------ */
/* SE3.__matmul__ (geometry.py:260-261).  round_f32 != 0 rounds the result to float32 as the
 * reference's float32 4x4 cache does. */
int vb_se3_compose_batch(const double* Ra, const double* ta, const double* Rb, const double* tb,
                         double* Rout, double* tout, int64_t n, int round_f32, void* stream);
/* SE3.inv (geometry.py:235-243): R^T, -R^T t; round_f32 != 0 mirrors the float32 store. */
int vb_se3_invert_batch(const double* R, const double* t, double* Rinv, double* tinv, int64_t n,
                        int round_f32, void* stream);
/* project_SO3 (geometry.py:175-191): U diag(1,1,det(U V^T)) V^T, one thread per block. */
int vb_polar_so3_batch(const double* M, double* R, int64_t n, void* stream);
/* The three SVD factors used by the primal/dual updates (bipgo.py:306-312, :323-329);
 * any output may be NULL. */
int vb_svd3_factors_batch(const double* M, double* rot, double* sym_pos, double* sym_inv, int64_t n,
                          void* stream);

/* ---- evaluation helpers (vican/geometry.py:131-172, :264-324; main.ipynb cell 9) ------------ */
/* optimize_gauge_SE3 (geometry.py:294-324): G = (project_SO3((sum_i Ra_i^T Rb_i)^T),
 * (1/n) sum_i Rb_i^T (ta_i - tb_i)); with ta == tb == NULL it is optimize_gauge_SO3
 * (geometry.py:264-291) and gauge_t is not written.  Deterministic two-stage reduction;
 * gauge_R [9], gauge_t [3] are device pointers. */
int64_t vb_gauge_workspace_bytes(int64_t n);
int vb_optimize_gauge(const double* Ra, const double* ta, const double* Rb, const double* tb, int64_t n,
                      double* gauge_R, double* gauge_t, void* workspace, int64_t workspace_bytes, void* stream);
/* distance_SO3 (geometry.py:154-172): angle(R1_i^T R2_i) in degrees, arccos of the clipped
 * trace as the reference; R2 == NULL gives angle(R1_i) (geometry.py:131-151). */
int vb_distance_so3_batch(const double* R1, const double* R2, double* deg, int64_t n, void* stream);
/* One transform applied from the left to n poses: (Rg, tg) @ (R_i, t_i)  (cell 9: G.inv() @ pose). */
int vb_se3_left_compose_batch(const double* Rg, const double* tg, const double* R, const double* t,
                              double* Rout, double* tout, int64_t n, int round_f32, void* stream);

/* ---- ingestion (bipgo.py:203-276, :420-431) ---------------------------------------------- */
/* Raw detections are given as flat arrays (already filtered by edge_filter on the host):
 * cam/time/marker indices, detection rotation R[E_raw][9], weights k_r, k_t.  markerC[m] =
 * R_m^T R_0 (constraint fold, bipgo.py:209-213).  round_kr_f32 mirrors numpy's float32
 * product k_r * R when the pose arrays are float32 (object variant, geometry.py:209-211).  */
int64_t vb_ingest_workspace_bytes(int64_t n_raw);
/* step 1: sort raw detections by (time, camera), count aggregated pairs.  Writes raw_perm
 * [n_raw] (sorted position -> raw index), raw_pair [n_raw] (pair id per sorted position) and
 * returns the number of pairs in *h_n_pairs (host; synchronises the stream). */
int vb_ingest_sort(const int32_t* cam, const int32_t* time, int64_t n_raw, int64_t n_c, int64_t n_t,
                   int32_t* raw_perm, int32_t* raw_pair, int64_t* h_n_pairs,
                   int32_t* h_sorted /* out, may be NULL: 1 if the detections arrived ordered by (time, camera) */,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* step 2: fold + aggregate into the time-sorted arrays, build the camera-pass copy, row
 * pointers, degrees and camera tiles.  pair_start [E+1] indexes the sorted raw list.  The
 * camera-pass arrays c_B / c_time / c_w are ordered by (time window, camera, time): c_order [E]
 * maps their positions to time-sorted edges, c_segptr [n_windows*n_c+1] delimits the run of every
 * (window, camera) -- per-camera reductions walk a camera's n_windows runs -- and the runs are cut
 * into tiles (runs of one camera, <= tile_len edges, contiguous: tile_start carries a
 * sentinel tile_start[n_tiles] = E; tile_off [n_windows*n_c+1] = first tile of every run).  tile_cam /
 * tile_start must hold vb_ingest_max_tiles + 1 entries.
 * PADDING: t_B / c_B must be allocated for E + 2 blocks and t_cam / c_time for E + 8 indices
 * (the edge passes stream them with 16-byte granular bulk copies).
 * WORKSPACE of vb_ingest_build: vb_ingest_workspace_bytes(max(n_raw, vb_ingest_windows(...) * n_c + 1)). */
/* Optional arrival schedule of the rotation array R of vb_ingest_build (host-to-device copy still in flight on
 * another stream): chunk k holds the detections [h_raw_end[k-1], h_raw_end[k]) and is complete when the CUDA event
 * h_events[k] (cudaEvent_t) has fired.  With sorted_input (vb_ingest_sort's h_sorted) every chunk is folded as soon as
 * it has landed, under the copy of the next one; otherwise the fold waits for all events.  n_chunks <= 64. */
typedef struct vb_arrival {
    int32_t n_chunks;
    int32_t sorted_input;
    const int64_t* h_raw_end;   /* [n_chunks] host */
    void* const*   h_events;    /* [n_chunks] host array of cudaEvent_t */
} vb_arrival;
int64_t vb_ingest_max_tiles(int64_t n_edges, int64_t n_c, int64_t tile_len);
int64_t vb_ingest_windows(int64_t n_edges, int64_t n_c, int64_t tile_len);
int vb_ingest_build(const int32_t* cam, const int32_t* time, const int32_t* marker, const double* R,
                    const double* k_r, const double* k_t, const double* markerC, int64_t n_raw,
                    int round_kr_f32, const int32_t* raw_perm, const int32_t* raw_pair, int64_t n_pairs,
                    int64_t n_c, int64_t n_t, int64_t tile_len,
                    int32_t* t_rowptr, int32_t* t_cam, int32_t* t_time, double* t_B, double* t_a, double* t_w,
                    int32_t* pair_start, int32_t* c_segptr, int32_t* c_time, double* c_B, double* c_w,
                    int32_t* c_order, int32_t* tile_cam, int32_t* tile_start, int32_t* tile_off,
                    int64_t* h_n_tiles, double* deg_t, double* deg_c,
                    const vb_arrival* arrival /* NULL: R is resident */,
                    int64_t n_markers /* rows of markerC */, int identity_perm /* raw_perm is the identity (h_sorted) */,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Number of connected components of the bipartite graph of aggregated edges (min-label hooking + pointer
 * jumping on the device).  labels: scratch [n_c + n_t + 2] int32; t_time [E] = time node of every time-sorted
 * edge.  The reference's early exit `max |lambda_1..5| <= 1e-6` (bipgo.py:283) can only fire with > 1 component.
 * Synchronises the stream. */
int vb_count_components(const vb_graph* g, const int32_t* t_time, int32_t* labels, int64_t* h_count, void* stream);

/* Incremental ingestion (streamed detections, cam.py:176-185, :243-263): a chunk of NEW time nodes is ingested
 * on its own (vb_ingest_sort / vb_ingest_build with chunk-local time indices) and appended behind the arrays
 * of the growing graph -- the time-sorted CSR is append-only in time, and the chunk's time windows become new
 * windows of the camera-pass order, so nothing already resident is re-sorted or rewritten.  These two helpers
 * do the index fix-ups: dst[i] = src[i] + add (row pointers, edge / tile / time offsets) and dst += src (camera
 * degrees).  Block and weight arrays are appended with plain device-to-device copies by the caller. */
int vb_offset_copy_i32(int32_t* dst, const int32_t* src, int64_t n, int32_t add, void* stream);
int vb_add_inplace_f64(double* dst, const double* src, int64_t n, void* stream);

/* ---- rotation stage (bipgo.py:243-348) --------------------------------------------------- */
/* One edge pass each (exposed for tests and for the roofline measurement).  Gathered node
 * blocks use the PADDED layout [n][S], S = vb_gather_stride() doubles (3 rows x 4 doubles, row = 32
 * bytes, then padding up to one 128-byte line) so that one row is one 256-bit load and the three
 * rows of a block share an L1 line; vb_pad_blocks converts a compact [n][9] array.
 *   vb_pass_time: out12_t = [Lambda_T[t]] * sum_{e in t} B_e^T X12[c_e]  (mode 0 with lamT [n_t][9], mode 1 raw sum)
 *   vb_pass_cam : Y_c     = sum_{e in c} B_e W12[t_e]      (Y compact [n_c][9]; per-tile sums go to
 *                 g->tile_part and are combined per camera in a fixed order; g->c_B holds the blocks
 *                 transposed, as vb_ingest_build writes them) */
int vb_gather_stride(void);
int vb_pad_blocks(const double* src9, double* dst12, int64_t n, void* stream);
int vb_pass_time(const vb_graph* g, int mode, const double* X12, const double* lamT, double* out12, void* stream);
int vb_pass_cam(const vb_graph* g, const double* W12, double* Y, void* stream);
/* Per-node updates: primal (bipgo.py:306-315): r_c, Lambda_C = U S U^T, Lambda_C^-1;
 * dual (bipgo.py:323-332): r_t, Lambda_T = U S^-1 U^T, and Wt12 = Lambda_T Y_t (Yt12 / Wt12 padded,
 * may alias). */
int vb_primal_update(const double* M, double* r_c, double* lamC, double* lamCinv, int64_t n_c, void* stream);
int vb_dual_update(const double* Yt12, double* r_t, double* lamT, double* Wt12, int64_t n_t, void* stream);
/* Gauge + projection (bipgo.py:295-297): X_c <- project_SO3(V_c V_0^-1). */
int vb_gauge_project(const double* V, double* r_c, int64_t n_c, void* stream);

int64_t vb_so3sync_workspace_bytes(int64_t n_c, int64_t n_t);
/* Whole primal-dual loop.  Outputs r_c [n_c][9], r_t [n_t][9] as stored by the reference
 * BEFORE its final transpose (bipgo.py:344-348): world rotations are the transposes.
 * Synchronises the stream (reads convergence flags).  stats may be NULL. */
int vb_so3sync_run(const vb_graph* g, const vb_so3_options* opt, double* r_c, double* r_t,
                   void* workspace, int64_t workspace_bytes, vb_so3_stats* stats, void* stream);

/* ---- translation stage (bipgo.py:420-487) ------------------------------------------------ */
/* Per aggregated pair g_p = sum_{raw e in p} k_t^2 d_e with
 *   d_e = Rw_c t_cm + Rw_t q_m,  q_m = (R_0^T R_m) (T_m^-1 T_0).t   (bipgo.py:451-455),
 * Rw = world rotations (transposes of r_c / r_t).  Also writes d_raw [n_raw][3] in sorted order
 * when non-NULL (needed by LSQR).  rhs = J^T t~ : rhs_c [n_c][3], rhs_t [n_t][3].  r_c_pad: scratch
 * [n_c][vb_gather_stride()] (the camera rotations are gathered from a padded copy with 256-bit loads).
 * n_markers = rows of marker_q (<= 256: staged in shared memory); identity_perm != 0: raw_perm is the identity
 * (vb_ingest_sort's h_sorted) and is not read. */
int vb_trans_rhs(const vb_graph* g, const int32_t* raw_perm, const int32_t* pair_start, const int32_t* marker,
                 const double* t_cm, const double* k_t, const double* marker_q, const double* r_c,
                 const double* r_t, const int32_t* t_time, double* pair_g,
                 double* d_sorted, double* rhs_c, double* rhs_t, double* r_c_pad, int64_t n_markers,
                 int identity_perm, void* stream);
/* Sliced-ELL copy of the translation Laplacian for vb_trans_cg (see vb_graph).  vb_sell_count writes the slice
 * pointers (st_ptr [ceil(n_t/8)+2], sc_ptr [ceil(n_c/8)+2]) and returns the chunk totals on the host
 * (synchronises); the caller allocates 32 * chunks slots per side and vb_sell_fill fills them. */
int64_t vb_sell_workspace_bytes(int64_t n_c, int64_t n_t, int64_t n_windows);
int vb_sell_count(const vb_graph* g, int32_t* st_ptr, int32_t* sc_ptr, int64_t* h_chunks_t, int64_t* h_chunks_c,
                  void* workspace, int64_t workspace_bytes, void* stream);
int vb_sell_fill(const vb_graph* g, const int32_t* st_ptr, int32_t* st_idx, double* st_w, const int32_t* sc_ptr,
                 int32_t* sc_idx, double* sc_w, int64_t chunks_c /* as returned by vb_sell_count */, void* workspace,
                 int64_t workspace_bytes, void* stream);
int64_t vb_trans_cg_workspace_bytes(int64_t n_c, int64_t n_t);
/* Conjugate gradients on J^T J x = J^T t~ replaying scipy.sparse.linalg.cg as the reference
 * calls it (bipgo.py:477: x0 = 0, no preconditioner, rtol = 1e-5, atol = 0, maxiter = 10 * 3N,
 * ||r|| tested before each step), including the ARITHMETIC of its CSR mat-vec: every row of J^T J p is the
 * left-to-right sum over ascending unknown index with separately rounded products (csrc/cg.cuh).
 * unk_c [n_c] / unk_t [n_t]: index of every camera / time node in the reference's unknown order (bipgo.py:
 * 420-430; it fixes where the diagonal entry sits inside a row); both NULL = cameras first, then time nodes.
 * Node indices must ascend with the unknown index inside each class.  Deterministic (no atomics): bitwise
 * reproducible run to run.  jacobi != 0 switches to the Jacobi-preconditioned variant (accurate mode; not
 * the reference iteration).  x_c [n_c][3], x_t [n_t][3].  Needs g->st_* / g->sc_*. */
int vb_trans_cg(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t,
                double rtol, int64_t maxiter, int jacobi, const int32_t* unk_c, const int32_t* unk_t,
                int32_t* h_iters, void* workspace,
                int64_t workspace_bytes, vb_allreduce_fn allreduce, void* allreduce_ctx,
                int owns_camera_diagonal /* 1 on a single GPU and on rank 0 of a sharded run */, void* stream);
int64_t vb_trans_lsqr_workspace_bytes(int64_t n_c, int64_t n_t, int64_t n_raw);
/* LSQR on J x = t~ replaying scipy.sparse.linalg.lsqr defaults (bipgo.py:480: damp 0,
 * atol = btol = 1e-6, conlim 1e8, iter_lim = 2 * 3N).  Rows are the raw detections in sorted
 * order: k_t (x_t - x_c) = k_t d_e.  allreduce != NULL (edge-sharded runs): rows and time nodes are local, the
 * camera block is replicated; three collectives per bidiagonalisation step (||u||^2; the camera part of A^T u;
 * ||v_t||^2 and ||w_t||^2 together). */
int vb_trans_lsqr(const vb_graph* g, const int32_t* raw_perm, const int32_t* raw_pair, const int32_t* pair_start,
                  const int32_t* t_time, const double* k_t, const double* d_sorted, int64_t n_raw, double* x_c, double* x_t,
                  double atol, double btol, double conlim, int64_t iter_lim, int32_t* h_istop,
                  int32_t* h_iters, void* workspace, int64_t workspace_bytes, vb_allreduce_fn allreduce,
                  void* allreduce_ctx, void* stream);

/* Dense direct solve of the same normal equations for small camera sets (n_c <= 8192; the
 * "direct" option in accurate mode): time nodes are eliminated in closed form, the n_c x n_c camera
 * Schur complement is factored by a blocked Cholesky written in this library (no cuSOLVER), and
 * the minimum-norm minimiser is returned -- the limit of the reference's cg / lsqr iterations
 * (bipgo.py:477-480), which stop ~1e-5 short of it.  Returns VB_STATUS_SINGULAR if the graph is
 * disconnected.  Synchronises the stream. */
int64_t vb_trans_schur_workspace_bytes(int64_t n_c, int64_t n_t);
int vb_trans_schur_direct(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t,
                          void* workspace, int64_t workspace_bytes, void* stream);

/* ---- multi-GPU (edge-sharded by time-node range, one process per GPU) ---------------------- */
/* NCCL is resolved at run time (dlopen); the reference has no distributed code, these exist so
 * that vb_so3sync_run / vb_trans_cg can sum camera-side accumulators across ranks. */
int vb_nccl_available(void);
int vb_nccl_unique_id(void* h_out128, int64_t bytes);                 /* rank 0, then broadcast */
int vb_nccl_init(const void* h_id128, int64_t bytes, int rank, int nranks, void** ctx_out);
int vb_nccl_destroy(void* ctx);
int vb_nccl_allreduce(void* ctx, double* buf, int64_t count, void* stream);   /* a vb_allreduce_fn */
void* vb_nccl_allreduce_fn(void);
/* One-shot all-reduce over NVLink peer memory (CUDA IPC windows, all ranks on one box, <= 8):
 * every rank sums the partials straight out of its peers' windows in rank order, so the result is
 * bitwise identical everywhere; one kernel, no staging copy.  vb_peer_create allocates this rank's
 * window (capacity_doubles per buffer; the only device allocation the library ever makes) and
 * returns its 64-byte IPC handle; after an all-gather of the handles vb_peer_connect maps the
 * others.  The caller must barrier between connect and first use and before destroy.
 * vb_peer_allreduce is a vb_allreduce_fn; with vb_so3_options.peer_ctx the exchange is the epilogue
 * of the camera-pass kernel itself. */
int vb_peer_create(int rank, int nranks, int64_t capacity_doubles, void** ctx_out, void* h_handle_out64);
int vb_peer_connect(void* ctx, const void* h_all_handles /* nranks x 64 bytes */);
int vb_peer_destroy(void* ctx);
int vb_peer_allreduce(void* ctx, double* buf, int64_t count, void* stream);
void* vb_peer_allreduce_fn(void);
/* Synchronises the stream; VB_STATUS_PEER_TIMEOUT if any wait on a peer gave up (~10 s) since creation. */
int vb_peer_status(void* ctx, void* stream);
/* Diagnostics: globaltimer stamps (ns) of the last exchange on this rank -- [0] publish (last CTA finished
 * its partial sums), [1] all flags seen by CTA 0, [2] CTA 0 finished its sums, [3] last CTA done.  Synchronises. */
int vb_peer_stamps(void* ctx, uint64_t* h_out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VICAN_B200_H */
