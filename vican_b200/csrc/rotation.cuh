// Rotation stage: per-node SVD kernels and the primal-dual loop driver
// (replaces vican/bipgo.py:271-348; one thread per node, registers only).
#pragma once
#include "../../include/vican_b200.h"
#include "common.cuh"
#include "lobpcg.cuh"
#include "passes.cuh"

namespace vb {

constexpr int NODE_THREADS = 128;

inline int node_grid(int64_t n) { return (int)((n + NODE_THREADS - 1) / NODE_THREADS); }

// bipgo.py:271-276: Lambda_T = I / deg_t, Lambda_C = pwr_deg I with pwr_deg_c = sum_t a_ct.
__global__ void init_lambda_kernel(const double* __restrict__ deg, double* __restrict__ lam, double* __restrict__ lam_inv,
                                   int64_t n, int invert) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = deg[i];
    const double v = invert ? 1.0 / d : d;
    double* L = lam + 9 * i;
    L[0] = v; L[1] = 0; L[2] = 0; L[3] = 0; L[4] = v; L[5] = 0; L[6] = 0; L[7] = 0; L[8] = v;
    if (lam_inv) {
        const double w = invert ? d : 1.0 / d;
        double* Li = lam_inv + 9 * i;
        Li[0] = w; Li[1] = 0; Li[2] = 0; Li[3] = 0; Li[4] = w; Li[5] = 0; Li[6] = 0; Li[7] = 0; Li[8] = w;
    }
}

__global__ void fill_identity_kernel(double* __restrict__ X, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* L = X + 9 * i;
    L[0] = 1; L[1] = 0; L[2] = 0; L[3] = 0; L[4] = 1; L[5] = 0; L[6] = 0; L[7] = 0; L[8] = 1;
}

// start block of the first eigen-solve, step 1: X = E_0 (identity at the gauge camera, zero elsewhere)
__global__ void fill_root_kernel(double* __restrict__ X, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = (i == 0) ? 1.0 : 0.0;
    double* L = X + 9 * i;
    L[0] = d; L[1] = 0; L[2] = 0; L[3] = 0; L[4] = d; L[5] = 0; L[6] = 0; L[7] = 0; L[8] = d;
}

// step 2: Y = P Lambda_T P^T E_0 holds, for every camera that shares a time node with the gauge camera, the
// weighted sum of its relative rotations to it: X_c = project_SO3(Y_c) is a one-hop spanning estimate of the
// solution (identity where nothing was seen).  The eigen-solve converges to the same invariant subspace from any
// start; from this one it needs about half the steps it needs from identity blocks.
__global__ void __launch_bounds__(NODE_THREADS) init_from_root_kernel(const double* __restrict__ Y, double* __restrict__ X, int64_t n_c) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    double m[9], rot[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = Y[9 * c + i];
    const double f2 = frob2_3(m);
    if (f2 > 1e-200 && isfinite(f2)) node_factors(m, rot, nullptr, nullptr);
    else { rot[0] = 1; rot[1] = 0; rot[2] = 0; rot[3] = 0; rot[4] = 1; rot[5] = 0; rot[6] = 0; rot[7] = 0; rot[8] = 1; }
#pragma unroll
    for (int i = 0; i < 9; ++i) X[9 * c + i] = rot[i];
}

// bipgo.py:295-297: r_c = project_SO3(V_c inv(V_0)); camera 0 is the gauge camera (first in the
// reference's lexicographic node order -- the host assigns indices in that order).
// r12 (optional): the same blocks in the padded gather layout (3 rows x 4 doubles) for the next time pass
__device__ __forceinline__ void store_padded(double* __restrict__ r12, int64_t c, const double* rot) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) r12[GSTRIDE * c + 4 * i + j] = rot[3 * i + j];
        r12[GSTRIDE * c + 4 * i + 3] = 0.0;
    }
}

__global__ void __launch_bounds__(NODE_THREADS) gauge_project_kernel(const double* __restrict__ V, double* __restrict__ r_c, int64_t n_c,
                                                                   double* __restrict__ r12 = nullptr) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    double v0[9], v0i[9], v[9], x[9], rot[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { v0[i] = V[i]; v[i] = V[9 * c + i]; }
    inv3(v0, v0i);
    mm3(v, v0i, x);
    node_factors(x, rot, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < 9; ++i) r_c[9 * c + i] = rot[i];
    if (r12) store_padded(r12, c, rot);
}

// Y_c <- Y_c R0^T for every camera block (R0: 9 doubles on the device).  Used when the eigen-iteration
// accepts its start block unchanged: see the shortcut in so3sync_run.
__global__ void __launch_bounds__(NODE_THREADS) rotate_right_transposed_kernel(double* __restrict__ Y, const double* __restrict__ R0, int64_t n_c,
                                                                               const double* __restrict__ skip_flag = nullptr) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;   // speculative launch (see so3sync_run)
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    double m[9], r0[9], o[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { m[i] = Y[9 * c + i]; r0[i] = R0[i]; }
    mmt3(m, r0, o);
#pragma unroll
    for (int i = 0; i < 9; ++i) Y[9 * c + i] = o[i];
}

// Deterministic pseudo-random start block of the second eigen-solve (any block with components outside the
// deflated subspace works; integer hash -> uniform in (-1, 1))
__global__ void fill_hash_kernel(double* __restrict__ X, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    X[i] = (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// Deflation of the converged eigenvectors V (orthonormal columns, [3 n_c][3]): the step kernel forms
// A S = Lambda_C S - Y, so  (L + shift V V^T) S  is obtained by  Y <- Y - shift V (V^T S).  One block.
__global__ void __launch_bounds__(1024) deflate_kernel(const double* __restrict__ V, const double* __restrict__ S, double* __restrict__ Y,
                                                      int64_t n_rows, double shift) {
    __shared__ double sm[32][9];
    __shared__ double G[9];
    double g[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) g[q] = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += 1024) {
        const double v0 = V[3 * r], v1 = V[3 * r + 1], v2 = V[3 * r + 2];
        const double s0 = S[3 * r], s1 = S[3 * r + 1], s2 = S[3 * r + 2];
        g[0] += v0 * s0; g[1] += v0 * s1; g[2] += v0 * s2;
        g[3] += v1 * s0; g[4] += v1 * s1; g[5] += v1 * s2;
        g[6] += v2 * s0; g[7] += v2 * s1; g[8] += v2 * s2;
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const double t = warp_sum(g[q]);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = t;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sm[w][threadIdx.x];
        G[threadIdx.x] = shift * t;
    }
    __syncthreads();
    for (int64_t r = threadIdx.x; r < n_rows; r += 1024) {
        const double v0 = V[3 * r], v1 = V[3 * r + 1], v2 = V[3 * r + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) Y[3 * r + j] -= v0 * G[j] + v1 * G[3 + j] + v2 * G[6 + j];
    }
}

// bipgo.py:306-315
__global__ void __launch_bounds__(NODE_THREADS) primal_update_kernel(const double* __restrict__ M, double* __restrict__ r_c, double* __restrict__ lamC,
                                     double* __restrict__ lamCinv, int64_t n_c, double* __restrict__ r12 = nullptr,
                                     const double* __restrict__ skip_flag = nullptr) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;   // speculative launch (see so3sync_run)
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_c) return;
    double m[9], rot[9], sp[9], si[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = M[9 * c + i];
    node_factors(m, rot, sp, si);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r_c[9 * c + i] = rot[i];
        lamC[9 * c + i] = sp[i];
        if (lamCinv) lamCinv[9 * c + i] = si[i];
    }
    if (r12) store_padded(r12, c, rot);
}

// bipgo.py:323-332 (+ Wt = Lambda_T Y_t, the time half of the next L-apply).  Yt12 / Wt12 use the
// padded gather layout (3 rows x 4 doubles) and may alias.  One thread per node; the padded rows move as
// 256-bit loads / stores and the two compact outputs (r_t, Lambda_T: 72-byte records) leave through a
// shared-memory transpose so that the CTA writes them as contiguous, coalesced streams.
__device__ __forceinline__ void st_row256(double* p, double a, double b, double c) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(0.0) : "memory");
}
__device__ __forceinline__ void ld_row256_plain(const double* p, double& a, double& b, double& c) {
    double pad;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(pad) : "l"(p));
    (void)pad;
}
__global__ void __launch_bounds__(NODE_THREADS) dual_update_kernel(const double* Yt12, double* __restrict__ r_t, double* __restrict__ lamT, double* Wt12,
                                   int64_t n_t, const double* __restrict__ skip_flag = nullptr) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;   // speculative launch (see so3sync_run); uniform over the grid
    __shared__ double sR[NODE_THREADS * 9], sL[NODE_THREADS * 9];
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x;
    const int64_t t = t0 + threadIdx.x;
    if (t < n_t) {
        double y[9], rot[9], si[9], w[9];
        ld_row256_plain(Yt12 + GSTRIDE * t, y[0], y[1], y[2]);
        ld_row256_plain(Yt12 + GSTRIDE * t + 4, y[3], y[4], y[5]);
        ld_row256_plain(Yt12 + GSTRIDE * t + 8, y[6], y[7], y[8]);
        node_factors(y, rot, nullptr, si);
        mm3(si, y, w);
#pragma unroll
        for (int i = 0; i < 9; ++i) { sR[9 * threadIdx.x + i] = rot[i]; sL[9 * threadIdx.x + i] = si[i]; }
        if (Wt12) {
            st_row256(Wt12 + GSTRIDE * t, w[0], w[1], w[2]);
            st_row256(Wt12 + GSTRIDE * t + 4, w[3], w[4], w[5]);
            st_row256(Wt12 + GSTRIDE * t + 8, w[6], w[7], w[8]);
        }
    }
    __syncthreads();
    const int64_t n_here = (n_t - t0) < NODE_THREADS ? (n_t - t0) : NODE_THREADS;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const int i = q * NODE_THREADS + threadIdx.x;
        if (i < 9 * n_here) { r_t[9 * t0 + i] = sR[i]; lamT[9 * t0 + i] = sL[i]; }
    }
}

__global__ void polar_batch_kernel(const double* __restrict__ M, double* __restrict__ R, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m[9], rot[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = M[9 * i + k];
    node_factors(m, rot, nullptr, nullptr);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[9 * i + k] = rot[k];
}

__global__ void svd_factors_batch_kernel(const double* __restrict__ M, double* rot, double* sp, double* si, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double m[9], a[9], b[9], c[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = M[9 * i + k];
    node_factors(m, a, b, c);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (rot) rot[9 * i + k] = a[k];
        if (sp) sp[9 * i + k] = b[k];
        if (si) si[9 * i + k] = c[k];
    }
}

// ----------------------------------------------------------------------------- workspace
inline int64_t align256(int64_t b) { return (b + 255) & ~(int64_t)255; }

struct So3Work {
    double *X, *AX, *W, *AW, *P, *AP, *Y, *Ykeep, *lamC, *lamCinv, *degc, *Xpad, *lamT, *Wt, *small, *partial;
    double *X2, *AX2, *small2;   // second eigen-solve (lambda_4..6 on the deflated operator, opt->eval_gap)
    int64_t bytes;
};

inline So3Work carve_so3(void* base, int64_t n_c, int64_t n_t) {
    So3Work w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t nd) {
        double* r = (double*)(p + off);
        off += align256(nd * (int64_t)sizeof(double));
        return r;
    };
    w.X = take(9 * n_c); w.AX = take(9 * n_c); w.W = take(9 * n_c); w.AW = take(9 * n_c);
    w.P = take(9 * n_c); w.AP = take(9 * n_c); w.Y = take(9 * n_c); w.Ykeep = take(9 * n_c);
    w.lamC = take(9 * n_c); w.lamCinv = take(9 * n_c); w.degc = take(n_c);
    w.Xpad = take(GSTRIDE * n_c);
    w.lamT = take(9 * n_t); w.Wt = take(GSTRIDE * n_t);
    w.small = take(SM_SIZE);
    w.partial = take(3 * (int64_t)1024 * LOB_NRED);
    w.X2 = take(9 * n_c); w.AX2 = take(9 * n_c); w.small2 = take(SM_SIZE);
    w.bytes = off;
    return w;
}

constexpr int STATUS_SLOTS = 4;
constexpr int PROF_EVENTS = 1024;   // begin/end pairs for up to 512 edge-pass launches per run
struct PinnedStatus {
    double* h = nullptr;
    cudaEvent_t ev[STATUS_SLOTS];
    cudaEvent_t prof[PROF_EVENTS];
    bool prof_ready = false;
    PinnedStatus() {
        cudaMallocHost((void**)&h, STATUS_SLOTS * SM_SIZE * sizeof(double));
        for (int i = 0; i < STATUS_SLOTS; ++i) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    }
    void ensure_prof() {
        if (prof_ready) return;
        for (int i = 0; i < PROF_EVENTS; ++i) cudaEventCreate(&prof[i]);
        prof_ready = true;
    }
};
inline PinnedStatus& pinned_state() {
    static thread_local PinnedStatus s;
    return s;
}
inline double* pinned_status() { return pinned_state().h; }

inline int so3sync_run(const vb_graph* g, const vb_so3_options* opt, double* r_c, double* r_t, void* workspace,
                       int64_t workspace_bytes, vb_so3_stats* stats, cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    if (n_c < 3 || opt->maxiter < 1) return VB_STATUS_BAD_ARGUMENT;
    So3Work w = carve_so3(workspace, n_c, n_t);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    vb_so3_stats local;
    vb_so3_stats* S = stats ? stats : &local;
    memset(S, 0, sizeof(*S));
    double* hs = pinned_status();
    const size_t cbytes = 9 * n_c * sizeof(double);
    int status = VB_STATUS_OK;

    // optional per-launch CUDA-event timing of the edge passes (bench.py's in-step roofline)
    PinnedStatus& pst = pinned_state();
    const bool prof = opt->profile_events != 0;
    if (prof) pst.ensure_prof();
    int n_prof = 0;                 // events used so far
    signed char prof_kind[PROF_EVENTS / 2];   // 0 = time pass, 1 = camera pass, -1 = skipped on the device
    int last_time_slot = -1, last_cam_slot = -1;
    auto prof_begin = [&](int kind) -> int {
        if (!prof || n_prof + 2 > PROF_EVENTS) return -1;
        const int slot = n_prof / 2;
        prof_kind[slot] = (signed char)kind;
        cudaEventRecord(pst.prof[n_prof], st);
        n_prof += 2;
        return slot;
    };
    auto prof_end = [&](int slot) { if (slot >= 0) cudaEventRecord(pst.prof[2 * slot + 1], st); };

    // prepadded: the producer of X (LOBPCG step, gauge / primal kernels) already wrote w.Xpad
    auto time_pass = [&](int mode, const double* X, double* out, const double* skip = nullptr, bool prepadded = false) -> int {
        S->time_passes++; S->kernel_launches += prepadded ? 1 : 2;
        int rc = prepadded ? 0 : launch_pad_blocks(X, w.Xpad, n_c, st, skip);   // gather source layout: 3 rows x 4 doubles
        if (rc) return rc;
        const int slot = prof_begin(0);
        rc = launch_pass_time(mode, g->t_rowptr, g->t_cam, g->t_B, w.Xpad, w.lamT, out, n_t, st, skip);
        prof_end(slot);
        last_time_slot = slot;
        return rc;
    };
    auto cam_pass = [&](const double* Wt, double* Y, const double* skip = nullptr) -> int {
        if (opt->peer_ctx != nullptr) {   // camera pass fused with its cross-rank sum (peer windows over NVLink)
            S->cam_passes++; S->kernel_launches++;
            const int slot = prof_begin(1);
            int rc = launch_pass_cam_fused(g->tile_cam, g->tile_start, g->c_time, g->c_B, Wt, Y, n_c, g->n_tiles,
                                           (PeerCtx*)opt->peer_ctx, st, skip);
            prof_end(slot);
            last_cam_slot = slot;
            return rc;
        }
        S->cam_passes++; S->kernel_launches += 2;
        const int slot = prof_begin(1);
        int rc = launch_pass_cam(g, Wt, Y, st, skip);   // per-tile sums + fixed-order combine: no zeroing, no atomics
        prof_end(slot);
        last_cam_slot = slot;
        if (rc) return rc;
        if (opt->allreduce) return opt->allreduce(opt->allreduce_ctx, Y, 9 * n_c, (void*)st);
        return 0;
    };
#define VB_RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

    if (n_t > 0) VB_CHECK(cudaMemsetAsync(w.Wt, 0, GSTRIDE * n_t * sizeof(double), st));   // pad lanes of the rows
    // initial duals (global camera degree on edge-sharded runs)
    VB_CHECK(cudaMemcpyAsync(w.degc, g->deg_c, n_c * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (opt->allreduce) VB_RC(opt->allreduce(opt->allreduce_ctx, w.degc, n_c, (void*)st));
    init_lambda_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.degc, w.lamC, w.lamCinv, n_c, 0);
    if (n_t > 0) init_lambda_kernel<<<node_grid(n_t), NODE_THREADS, 0, st>>>(g->deg_t, w.lamT, nullptr, n_t, 1);
    fill_identity_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.X, n_c);
    VB_KERNEL_CHECK();
    S->kernel_launches += 3;

    LobpcgParams lp;
    lp.n_c = (int)n_c;
    lp.X = w.X; lp.AX = w.AX; lp.W = w.W; lp.AW = w.AW; lp.P = w.P; lp.AP = w.AP;
    lp.Y = w.Y; lp.lamC = w.lamC; lp.lamCinv = w.lamCinv; lp.small = w.small; lp.partial = w.partial;
    lp.tol = opt->tol;
    lp.Wpad = w.Xpad;
    const int max_inner = opt->max_inner > 0 ? opt->max_inner : 200;
    static const bool lob_timing = getenv("VICAN_B200_LOBPCG_TIMING") != nullptr;   // diagnostics: stage times of every step
    // VICAN_B200_SPEC=0: always speculate a second eigen-step, never the converged continuation (A/B switch)
    static const bool spec_continue = !(getenv("VICAN_B200_SPEC") && atoi(getenv("VICAN_B200_SPEC")) == 0);

    double max_eval = 1.0;   // bipgo.py:280
    int last_inner[2] = {0, 0};   // steps of the last two outer iterations (inexact-inner verification)
    bool used_early = false;
    for (int outer = 0; outer < opt->maxiter; ++outer) {
        if (opt->eval_gap && max_eval <= 1e-6) { S->early_exit = 1; break; }   // bipgo.py:283-284
        const bool early = opt->tol_early > 0.0 && opt->maxiter - outer > opt->early_margin;
        lp.tol = early ? opt->tol_early : opt->tol;
        used_early = used_early || early;
        if (outer == 0) {
            if (opt->identity_start == 0) {   // one-hop spanning start (2 extra edge passes, see init_from_root_kernel)
                fill_root_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.X, n_c);
                VB_RC(time_pass(0, w.X, w.Wt));
                VB_RC(cam_pass(w.Wt, w.Y));
                init_from_root_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.Y, w.X, n_c);
                VB_KERNEL_CHECK();
                S->kernel_launches += 2;
            }
            VB_RC(time_pass(0, w.X, w.Wt));
        }
        VB_RC(cam_pass(w.Wt, w.Y));
        // A host-issued collective (NCCL hook without the fused peer path) also runs for the speculative camera
        // pass that follows a converged step and would sum the already summed Y once more: keep a copy for the
        // shortcut below.  (The fused kernel and the single-GPU path leave Y alone when the flag is set.)
        const bool keep_Y = opt->allreduce != nullptr && opt->peer_ctx == nullptr;
        if (keep_Y) VB_CHECK(cudaMemcpyAsync(w.Ykeep, w.Y, cbytes, cudaMemcpyDeviceToDevice, st));
        // W, AW, P, AP are carved back to back: one memset
        VB_CHECK(cudaMemsetAsync(w.W, 0, (size_t)((char*)w.AP - (char*)w.W) + cbytes, st));
        lp.first = 1;
        VB_RC(launch_lobpcg_step(lp, st));
        S->lobpcg_steps++; S->kernel_launches++;
        // The host polls the convergence flag ONE step late: step k+1 (edge passes + LOBPCG kernel)
        // is enqueued before the status of step k is read, so the GPU never idles on the host.  Once
        // the flag is set on the device the speculative launches return immediately.
        PinnedStatus& ps = pinned_state();
        const double* conv_flag = w.small + SM_CONV;
        auto readback = [&](int step) -> int {
            const int slot = step % STATUS_SLOTS;
            VB_CHECK(cudaMemcpyAsync(ps.h + slot * SM_SIZE, w.small, SM_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
            VB_CHECK(cudaEventRecord(ps.ev[slot], st));
            return 0;
        };
        VB_RC(readback(1));
        int inner = 1;            // steps whose work was (or will be) really executed
        int enqueued = 1;         // steps enqueued so far
        double hist[3] = {1e300, 1e300, 1e300};   // residuals of the last three steps (stagnation guard)
        // Once the outer iteration has settled, the eigen-iteration accepts its start block at the first step
        // (3, 1, 1, ... steps per outer iteration).  After such an iteration the host no longer speculates a SECOND
        // step but the converged CONTINUATION: the shortcut's primal multiply, the primal update, the raw time pass
        // and the dual update are enqueued behind the first step with the step's not-converged flag as their skip
        // flag, and only then is the step's status read.  The GPU runs straight through (no queue of no-op launches,
        // no idle time while the host reacts); if the prediction fails the four launches returned at their first
        // instruction and the regular loop below takes over.  No collective is ever speculated.
        bool by_prediction = false;
        if (spec_continue && opt->no_shortcut == 0 && outer >= 1 && last_inner[1] == 1 && !opt->eval_gap && !keep_Y) {
            const double* nconv = w.small + SM_NCONV;
            rotate_right_transposed_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.Y, r_c, n_c, nconv);
            primal_update_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.Y, r_c, w.lamC, w.lamCinv, n_c, w.Xpad, nconv);
            VB_KERNEL_CHECK();
            VB_RC(time_pass(1, r_c, w.Wt, nconv, true));
            if (n_t > 0) dual_update_kernel<<<node_grid(n_t), NODE_THREADS, 0, st>>>(w.Wt, r_t, w.lamT, w.Wt, n_t, nconv);
            VB_KERNEL_CHECK();
            VB_CHECK(cudaEventSynchronize(ps.ev[1 % STATUS_SLOTS]));
            hs = ps.h + (1 % STATUS_SLOTS) * SM_SIZE;
            if (hs[SM_CONV] != 0.0) {
                by_prediction = true;
                S->kernel_launches += 3;
                S->shortcut_outer++;
            } else {   // skipped on the device
                S->time_passes--; S->kernel_launches--;
                if (last_time_slot >= 0) prof_kind[last_time_slot] = -1;
            }
        }
        while (!by_prediction) {
            bool speculated = false;
            if (enqueued < max_inner) {
                VB_RC(time_pass(0, w.W, w.Wt, conv_flag, true));   // the step kernel wrote W into Xpad as well
                VB_RC(cam_pass(w.Wt, w.Y, conv_flag));
                lp.first = 0;
                VB_RC(launch_lobpcg_step(lp, st));
                S->lobpcg_steps++; S->kernel_launches++;
                ++enqueued;
                VB_RC(readback(enqueued));
                speculated = true;
            }
            const int slot = inner % STATUS_SLOTS;
            VB_CHECK(cudaEventSynchronize(ps.ev[slot]));
            hs = ps.h + slot * SM_SIZE;
            if (lob_timing) {
                fprintf(stderr, "[lobpcg] outer %d step %d conv %.0f: stage1 %.1f sync1 %.1f combine %.1f ritz %.1f stage3 %.1f sync2+c %.1f stage4 %.1f sync3+c %.1f stage5 %.1f total %.1f us, %.0f Jacobi sweeps\n",
                        outer, inner, hs[SM_CONV], 1e-3 * (hs[SM_TIME + 1] - hs[SM_TIME]), 1e-3 * (hs[SM_TIME + 2] - hs[SM_TIME + 1]),
                        1e-3 * (hs[SM_TIME + 3] - hs[SM_TIME + 2]), 1e-3 * (hs[SM_TIME + 4] - hs[SM_TIME + 3]),
                        1e-3 * (hs[SM_TIME + 5] - hs[SM_TIME + 4]), 1e-3 * (hs[SM_TIME + 6] - hs[SM_TIME + 5]),
                        1e-3 * (hs[SM_TIME + 7] - hs[SM_TIME + 6]), 1e-3 * (hs[SM_TIME + 8] - hs[SM_TIME + 7]),
                        1e-3 * (hs[SM_TIME + 9] - hs[SM_TIME + 8]), 1e-3 * (hs[SM_TIME + 9] - hs[SM_TIME]), hs[SM_TIME + 10]);
            }
            if (hs[SM_CONV] != 0.0) {
                if (speculated) {
                    S->time_passes--; S->cam_passes--; S->lobpcg_steps--; S->kernel_launches -= (opt->peer_ctx != nullptr) ? 3 : 4;
                    if (last_time_slot >= 0) prof_kind[last_time_slot] = -1;
                    if (last_cam_slot >= 0) prof_kind[last_cam_slot] = -1;
                }
                break;
            }
            if (!speculated) { S->stalled_outer++; status = VB_STATUS_EIG_STALLED; break; }
            // rounding floor reached: the residual stopped shrinking although it is already tiny.  The
            // step enqueued above still runs (its flag test sees "not converged"); take its iterate.
            const double rmax = fmax(hs[SM_RESN], fmax(hs[SM_RESN + 1], hs[SM_RESN + 2]));
            const bool floor_hit = (rmax < 1e-9 * hs[SM_ANORM] && rmax > 0.5 * hist[0]);
            hist[0] = hist[1]; hist[1] = hist[2]; hist[2] = rmax;
            ++inner;
            if (floor_hit) {
                VB_CHECK(cudaEventSynchronize(ps.ev[inner % STATUS_SLOTS]));
                hs = ps.h + (inner % STATUS_SLOTS) * SM_SIZE;
                break;
            }
        }
        if (outer < 64) S->inner_per_outer[outer] = inner;
        last_inner[0] = last_inner[1]; last_inner[1] = inner;
        for (int j = 0; j < 3; ++j) { S->theta[j] = hs[SM_THETA + j]; S->resid[j] = hs[SM_RESN + j]; }
        S->anorm = hs[SM_ANORM];

        // Diagnostics of the reference (bipgo.py:288-292): the two eigenvalues after the wanted three, by the same
        // eigen-iteration on L + shift V V^T (V = the converged eigenvectors, moved out of the way); plain loop
        // with a host check per step -- only runs on request (verbose callers, graphs that are not connected).
        double hcopy[SM_SIZE];
        if (opt->eval_gap) {
            memcpy(hcopy, hs, sizeof(hcopy));   // the status ring is reused by the read-backs below
            hs = hcopy;
            double th3[3] = {hs[SM_THETA], hs[SM_THETA + 1], hs[SM_THETA + 2]};
            const double shift = 2.0 * hs[SM_ANORM];
            VB_CHECK(cudaMemcpyAsync(w.Ykeep, w.Y, cbytes, cudaMemcpyDeviceToDevice, st));
            VB_CHECK(cudaMemsetAsync(w.W, 0, (size_t)((char*)w.AP - (char*)w.W) + cbytes, st));
            fill_hash_kernel<<<node_grid(9 * n_c), NODE_THREADS, 0, st>>>(w.X2, 9 * n_c);
            LobpcgParams l2 = lp;
            l2.X = w.X2; l2.AX = w.AX2; l2.small = w.small2;
            double lam[3] = {0.0, 0.0, 0.0};
            for (int stp = 0; stp < max_inner; ++stp) {
                const double* blk = stp == 0 ? w.X2 : w.W;
                VB_RC(time_pass(0, blk, w.Wt, nullptr, stp != 0));
                VB_RC(cam_pass(w.Wt, w.Y));
                deflate_kernel<<<1, 1024, 0, st>>>(w.X, blk, w.Y, 3 * n_c, shift);
                l2.first = stp == 0 ? 1 : 0;
                VB_RC(launch_lobpcg_step(l2, st));
                S->kernel_launches += 2;
                VB_CHECK(cudaMemcpyAsync(pst.h, w.small2, SM_SIZE * sizeof(double), cudaMemcpyDeviceToHost, st));
                VB_CHECK(cudaStreamSynchronize(st));
                for (int j = 0; j < 3; ++j) lam[j] = pst.h[SM_THETA + j];
                const double rmax = fmax(pst.h[SM_RESN], fmax(pst.h[SM_RESN + 1], pst.h[SM_RESN + 2]));
                if (pst.h[SM_CONV] != 0.0 || rmax < 1e-9 * pst.h[SM_ANORM]) break;   // diagnostics: 1e-9 relative is plenty
            }
            // the wanted triple may hold the shifted copies when fewer than three further eigenvalues lie below the shift
            double ev5[5] = {th3[0], th3[1], th3[2], lam[0], lam[1]};
            max_eval = 0.0;
            for (int j = 0; j < 5; ++j) max_eval = fmax(max_eval, fabs(ev5[j]));   // bipgo.py:292
            if (outer < 64) for (int j = 0; j < 5; ++j) S->evals_hist[outer][j] = ev5[j];
            VB_CHECK(cudaMemcpyAsync(w.Y, w.Ykeep, cbytes, cudaMemcpyDeviceToDevice, st));
        }

        // Primal multiply M = P Lambda_T P^T r_c with r_c = project_SO3(V_c V_0^-1) (bipgo.py:295-300).
        // Shortcut: when the eigen-iteration accepted its start block at the FIRST step (inner == 1), the new
        // eigenvectors are V = R C with R = the previous r_c (blocks in SO(3)) and C a 3x3 matrix, so
        // V_c V_0^-1 = R_c C C^-1 R_0^-1 = R_c R_0^T is already a rotation, r_c = R R_0^T, and
        // M = (P Lambda_T P^T R) R_0^T = Y R_0^T with Y the camera-pass result this very iteration started
        // from: no gauge kernel, no time pass, no camera pass (2 instead of 4 edge passes per converged
        // outer iteration).  Y and the old r_c are still intact here (speculative passes leave Y alone).
        const bool shortcut = opt->no_shortcut == 0 && outer >= 1 && inner == 1 && hs[SM_CONV] != 0.0;
        if (by_prediction) {
            // the continuation already ran behind the first step (see above)
        } else if (shortcut) {
            if (keep_Y) VB_CHECK(cudaMemcpyAsync(w.Y, w.Ykeep, cbytes, cudaMemcpyDeviceToDevice, st));
            rotate_right_transposed_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.Y, r_c, n_c);
            VB_KERNEL_CHECK();
            S->shortcut_outer++;
        } else {
            gauge_project_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.X, r_c, n_c, w.Xpad);
            VB_KERNEL_CHECK();
            VB_RC(time_pass(0, r_c, w.Wt, nullptr, true));
            VB_RC(cam_pass(w.Wt, w.Y));
        }
        if (!by_prediction) {
            primal_update_kernel<<<node_grid(n_c), NODE_THREADS, 0, st>>>(w.Y, r_c, w.lamC, w.lamCinv, n_c, w.Xpad);
            VB_KERNEL_CHECK();
            VB_RC(time_pass(1, r_c, w.Wt, nullptr, true));
            if (n_t > 0) dual_update_kernel<<<node_grid(n_t), NODE_THREADS, 0, st>>>(w.Wt, r_t, w.lamT, w.Wt, n_t);
            VB_KERNEL_CHECK();
            S->kernel_launches += 3;
        }
        VB_CHECK(cudaMemcpyAsync(w.X, r_c, cbytes, cudaMemcpyDeviceToDevice, st));
        S->outer_done = outer + 1;
    }
    VB_CHECK(cudaStreamSynchronize(st));
    // inexact early iterations are only trusted when the tight end-game shows the outer iteration AT its fixed point
    if (used_early && !S->early_exit && !(last_inner[0] == 1 && last_inner[1] == 1)) S->inexact_unverified = 1;
    if (prof) {
        for (int slot = 0; slot < n_prof / 2; ++slot) {
            if (prof_kind[slot] < 0) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, pst.prof[2 * slot], pst.prof[2 * slot + 1]) != cudaSuccess) continue;
            if (prof_kind[slot] == 0) { S->time_pass_ms += ms; S->time_pass_timed++; }
            else { S->cam_pass_ms += ms; S->cam_pass_timed++; }
        }
    }
    return status;
#undef VB_RC
}

}  // namespace vb
