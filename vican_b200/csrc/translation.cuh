// Translation stage (replaces vican/bipgo.py:434-481).
//
// The reference builds the 3E_raw x 3N incidence matrix J (rows k_t (x_t - x_c)) and solves
// with scipy's cg on J^T J or lsqr on J.  Here J is never formed:
//   * J^T J = (bipartite graph Laplacian with weights w_ct = sum k_t^2) (x) I_3, applied in two
//     gather passes over the aggregated pairs (time-sorted, camera-sorted);
//   * the iterations replay scipy's recurrences and stopping rules exactly (SURVEY.md appendix
//     A), because the reference result is a TRUNCATED iterate, not the exact minimiser;
//   * all scalars (rho, alpha, beta, norms, stop flags) live on the device.
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
// thread per pair: g_p = sum k_t^2 d_e,  d_e = r_c^T t_cm + r_t^T q_m   (world rotation = r^T)
__global__ void trans_pair_kernel(const int* __restrict__ raw_perm, const int* __restrict__ pair_start,
                                  const int* __restrict__ marker, const double* __restrict__ t_cm,
                                  const double* __restrict__ k_t, const double* __restrict__ marker_q,
                                  const double* __restrict__ r_c, const double* __restrict__ r_t,
                                  const int* __restrict__ t_cam, const int* __restrict__ t_time, int64_t n_pairs,
                                  double* __restrict__ pair_g, double* __restrict__ d_sorted) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double Rc[9], Rt[9];
    const int64_t c = t_cam[p], t = t_time[p];
#pragma unroll
    for (int i = 0; i < 9; ++i) { Rc[i] = r_c[9 * c + i]; Rt[i] = r_t[9 * t + i]; }
    double g0 = 0, g1 = 0, g2 = 0;
    for (int pos = pair_start[p]; pos < pair_start[p + 1]; ++pos) {
        const int64_t r = raw_perm[pos];
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

// ------------------------------------------------------------------------------------- CG
// device scalars
enum { CG_RHO = 0, CG_RHO_NEXT, CG_PQ_T, CG_PQ_C, CG_RN2_T, CG_RN2_C, CG_BN2, CG_ATOL2, CG_ALPHA, CG_BETA,
       CG_DONE, CG_ITERS, CG_RHO_NEXT_T, CG_RHO_NEXT_C, CG_NSCAL = 16 };

struct CgWork {
    double *r_c, *p_c, *q_c, *dg_c;   // [3 n_c (+8 for packed scalars on q_c)]
    double *r_t, *p_t, *q_t, *dg_t;   // [3 n_t]
    double* sc;                        // [CG_NSCAL]
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
    // search directions are kept PADDED ([n][4], 32-byte rows) so the matvec gathers them with one 256-bit load
    w.r_c = take(3 * n_c); w.p_c = take(4 * n_c); w.q_c = take(3 * n_c + 8); w.dg_c = take(n_c);
    w.r_t = take(3 * n_t); w.p_t = take(4 * n_t); w.q_t = take(3 * n_t); w.dg_t = take(n_t);
    w.sc = take(CG_NSCAL);
    w.bytes = off;
    return w;
}

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

// r = b, x = 0, rho_next = r.z, rn2 = r.r (z = r or r/diag)
__global__ void cg_init_kernel(const double* __restrict__ b, const double* __restrict__ dg, int jacobi, double* __restrict__ x,
                               double* __restrict__ r, double* __restrict__ p, int64_t n_nodes, double* sc, int rho_slot, int rn_slot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (i < n_nodes) {
        const double d = jacobi ? 1.0 / dg[i] : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double bv = b[3 * i + k];
            x[3 * i + k] = 0.0; r[3 * i + k] = bv; p[4 * i + k] = 0.0;
            v[0] += bv * bv * d; v[1] += bv * bv;
        }
        p[4 * i + 3] = 0.0;
    }
    double* const dst[2] = {sc + rho_slot, sc + rn_slot};
    block_atomic_sum<2>(v, dst);
}

// one thread: top-of-iteration test + beta (scipy cg: `if norm(r) < atol: return`)
__global__ void cg_scalar_top_kernel(double* sc, double rtol, int first) {
    if (first) {
        sc[CG_BN2] = sc[CG_RN2_T] + sc[CG_RN2_C];
        const double atol = rtol * sqrt(sc[CG_BN2]);
        sc[CG_ATOL2] = atol;
        sc[CG_RHO] = 0.0;
    }
    if (sc[CG_DONE] != 0.0) return;
    const double rn = sqrt(sc[CG_RN2_T] + sc[CG_RN2_C]);
    if (rn < sc[CG_ATOL2]) { sc[CG_DONE] = 1.0; return; }
    const double rho = sc[CG_RHO_NEXT_T] + sc[CG_RHO_NEXT_C];
    sc[CG_BETA] = first ? 0.0 : rho / sc[CG_RHO];
    sc[CG_RHO] = rho;
    sc[CG_PQ_T] = 0.0; sc[CG_PQ_C] = 0.0;
    sc[CG_ITERS] += 1.0;
}

// p = z + beta p   (z = r or r / diag)
__global__ void cg_dir_kernel(const double* __restrict__ r, const double* __restrict__ dg, int jacobi, double* __restrict__ p,
                              int64_t n_nodes, const double* sc) {
    if (sc[CG_DONE] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const double beta = sc[CG_BETA];
    const double d = jacobi ? 1.0 / dg[i] : 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) p[4 * i + k] = r[3 * i + k] * d + beta * p[4 * i + k];
}

// time side of q = (J^T J) p : warp per time node (persistent, grid-stride); also accumulates p_t . q_t.
// Software pipelined: the indices / weights of the warp's NEXT node (two edges per lane, i.e. up to 64
// edges) and the row pointers of the one after are in flight while the current node's rows are gathered
// and reduced, so the only exposed latency per node is the (L2-resident) gather.
// The product is evaluated as  dg_t p_t - sum w p_c  (diagonal term separate, like the explicit CSR
// product of J^T J that scipy's cg multiplies with): measured on the object-calibration graphs, the
// truncated CG iterate is ~8x less sensitive to this rounding pattern than to sum w (p_t - p_c).
__global__ void __launch_bounds__(TR_THREADS)
cg_time_kernel(const int* __restrict__ rowptr, const int* __restrict__ cam, const double* __restrict__ w,
               const double* __restrict__ dg_t, const double* __restrict__ p_c, const double* __restrict__ p_t,
               double* __restrict__ q_t, int64_t n_t, double* sc) {
    if (sc[CG_DONE] != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double dot[1] = {0.0};
    auto load_head = [&](int s, int e, int& ca, int& cb, double& wa, double& wb) {
        const int i = s + lane, i2 = i + 32;
        ca = 0; cb = 0; wa = 0.0; wb = 0.0;
        if (i < e) { ca = cam[i]; wa = w[i]; }
        if (i2 < e) { cb = cam[i2]; wb = w[i2]; }
    };
    int s = 0, e = 0, s2 = 0, e2 = 0;
    if (warp0 < n_t) { s = __ldg(rowptr + warp0); e = __ldg(rowptr + warp0 + 1); }
    if (warp0 + nwarps < n_t) { s2 = __ldg(rowptr + warp0 + nwarps); e2 = __ldg(rowptr + warp0 + nwarps + 1); }
    int ca, cb; double wa, wb;
    load_head(s, e, ca, cb, wa, wb);
    for (int64_t node = warp0; node < n_t; node += nwarps) {
        // prefetch: head of the next node, row pointers of the one after
        int nca, ncb; double nwa, nwb;
        load_head(s2, e2, nca, ncb, nwa, nwb);
        const int64_t n3 = node + 2 * nwarps;
        int s3 = 0, e3 = 0;
        if (n3 < n_t) { s3 = __ldg(rowptr + n3); e3 = __ldg(rowptr + n3 + 1); }
        const double x0 = p_t[4 * node], x1 = p_t[4 * node + 1], x2 = p_t[4 * node + 2];
        double a0 = 0, a1 = 0, a2 = 0;
        {   // first 64 edges: indices / weights already in registers
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, h0 = 0.0, h1 = 0.0, h2 = 0.0;
            if (s + lane < e) ld_row256(p_c + 4 * (int64_t)ca, g0, g1, g2);
            if (s + lane + 32 < e) ld_row256(p_c + 4 * (int64_t)cb, h0, h1, h2);
            a0 = wa * g0 + wb * h0; a1 = wa * g1 + wb * h1; a2 = wa * g2 + wb * h2;
        }
        for (int i = s + 64 + lane; i < e; i += 32) {   // high-degree tail
            const int64_t c = cam[i];
            const double ww = w[i];
            double g0, g1, g2;
            ld_row256(p_c + 4 * c, g0, g1, g2);
            a0 += ww * g0; a1 += ww * g1; a2 += ww * g2;
        }
        {   // three sums with a reduce-scatter butterfly (6 + 2 instead of 15 fp64 shuffles: the kernel is bound by
            // the L1 data pipe, which the shuffles share with the row gathers): the totals land in lanes 0, 8, 16
            const double v3[3] = {a0, a1, a2};
            double mine = 0.0;
            RsStep<3, 16>::run(v3, lane, 0, 3, [&](int, double t) { mine = t; });
            a0 = mine;
            a1 = shfl(mine, 8);
            a2 = shfl(mine, 16);
        }
        if (lane == 0) {
            const double d = dg_t[node];
            a0 = d * x0 - a0; a1 = d * x1 - a1; a2 = d * x2 - a2;
            q_t[3 * node] = a0; q_t[3 * node + 1] = a1; q_t[3 * node + 2] = a2;
            dot[0] += x0 * a0 + x1 * a1 + x2 * a2;
        }
        s = s2; e = e2; s2 = s3; e2 = e3;
        ca = nca; cb = ncb; wa = nwa; wb = nwb;
    }
    double* const dst[1] = {sc + CG_PQ_T};
    block_atomic_sum<1>(dot, dst);
}

// camera side: warp per camera tile, accumulates -sum w p_t with 3 atomics per tile (q_c holds dg_c p_c);
// the indices / weights of the next 64-edge chunk are loaded before the current chunk's rows are gathered
__global__ void cg_cam_kernel(const int* __restrict__ tile_cam, const int* __restrict__ tile_start, const int* __restrict__ tile_end,
                              const int* __restrict__ tidx, const double* __restrict__ w, const double* __restrict__ p_c,
                              const double* __restrict__ p_t, double* __restrict__ q_c, int64_t n_tiles, const double* sc) {
    if (sc[CG_DONE] != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n_tiles) return;
    const int64_t c = tile_cam[warp];
    const int ts = tile_start[warp], te = tile_end[warp];
    auto load_chunk = [&](int base, int& ta, int& tb, double& wa, double& wb) {
        const int i = base + lane, i2 = i + 32;
        ta = 0; tb = 0; wa = 0.0; wb = 0.0;
        if (i < te) { ta = tidx[i]; wa = w[i]; }
        if (i2 < te) { tb = tidx[i2]; wb = w[i2]; }
    };
    double a0 = 0, a1 = 0, a2 = 0;
    int ta, tb; double wa, wb;
    load_chunk(ts, ta, tb, wa, wb);
    for (int base = ts; base < te; base += 64) {
        int nta, ntb; double nwa, nwb;
        load_chunk(base + 64, nta, ntb, nwa, nwb);
        double g0 = 0.0, g1 = 0.0, g2 = 0.0, h0 = 0.0, h1 = 0.0, h2 = 0.0;
        if (base + lane < te) ld_row256(p_t + 4 * (int64_t)ta, g0, g1, g2);
        if (base + lane + 32 < te) ld_row256(p_t + 4 * (int64_t)tb, h0, h1, h2);
        a0 -= wa * g0 + wb * h0;
        a1 -= wa * g1 + wb * h1;
        a2 -= wa * g2 + wb * h2;
        ta = nta; tb = ntb; wa = nwa; wb = nwb;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { atomicAdd(q_c + 3 * c, a0); atomicAdd(q_c + 3 * c + 1, a1); atomicAdd(q_c + 3 * c + 2, a2); }
}

// q_c = dg_c p_c on the rank that owns the diagonal term (rank 0 of an edge-sharded run: the
// camera accumulators are summed over ranks afterwards), zero elsewhere; also clears the 8 pack slots
__global__ void cg_qc_init_kernel(const double* __restrict__ dg_c, const double* __restrict__ p_c, double* __restrict__ q_c,
                                  int64_t n_c, int owner, const double* sc) {
    if (sc[CG_DONE] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_c) {
        const double d = owner ? dg_c[i] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) q_c[3 * i + k] = d * p_c[4 * i + k];
    }
    if (i < 8) q_c[3 * n_c + i] = 0.0;
}

// p_c . q_c (after the camera pass / all-reduce); the time part may have been packed at q_c[3 n_c]
__global__ void cg_dot_kernel(const double* __restrict__ p4, const double* __restrict__ q3, int64_t n_nodes, double* sc, int slot) {
    if (sc[CG_DONE] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (i < n_nodes) v[0] = p4[4 * i] * q3[3 * i] + p4[4 * i + 1] * q3[3 * i + 1] + p4[4 * i + 2] * q3[3 * i + 2];
    double* const dst[1] = {sc + slot};
    block_atomic_sum<1>(v, dst);
}

__global__ void cg_scalar_alpha_kernel(double* sc) {
    if (sc[CG_DONE] != 0.0) return;
    sc[CG_ALPHA] = sc[CG_RHO] / (sc[CG_PQ_T] + sc[CG_PQ_C]);
    sc[CG_RN2_T] = 0.0; sc[CG_RN2_C] = 0.0; sc[CG_RHO_NEXT_T] = 0.0; sc[CG_RHO_NEXT_C] = 0.0;
}

// x += alpha p; r -= alpha q; accumulate r.r and r.z
__global__ void cg_update_kernel(const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dg, int jacobi,
                                 double* __restrict__ x, double* __restrict__ r, int64_t n_nodes, double* sc, int rho_slot,
                                 int rn_slot) {
    if (sc[CG_DONE] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (i < n_nodes) {
        const double alpha = sc[CG_ALPHA];
        const double d = jacobi ? 1.0 / dg[i] : 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[3 * i + k] += alpha * p[4 * i + k];
            const double rv = r[3 * i + k] - alpha * q[3 * i + k];
            r[3 * i + k] = rv;
            v[0] += rv * rv * d; v[1] += rv * rv;
        }
    }
    double* const dst[2] = {sc + rho_slot, sc + rn_slot};
    block_atomic_sum<2>(v, dst);
}

__global__ void pack_scalars_kernel(double* dst, const double* sc, int s0, int s1, int s2) {
    dst[0] = s0 >= 0 ? sc[s0] : 0.0; dst[1] = s1 >= 0 ? sc[s1] : 0.0; dst[2] = s2 >= 0 ? sc[s2] : 0.0;
}
__global__ void unpack_scalars_kernel(const double* src, double* sc, int s0, int s1, int s2) {
    if (s0 >= 0) sc[s0] = src[0];
    if (s1 >= 0) sc[s1] = src[1];
    if (s2 >= 0) sc[s2] = src[2];
}

inline int trans_cg(const vb_graph* g, const double* rhs_c, const double* rhs_t, double* x_c, double* x_t, double rtol,
                    int64_t maxiter, int jacobi, int32_t* h_iters, void* workspace, int64_t workspace_bytes,
                    vb_allreduce_fn allreduce, void* actx, int owner, cudaStream_t st) {
    const int64_t n_c = g->n_c, n_t = g->n_t;
    CgWork w = carve_cg(workspace, n_c, n_t);
    if (w.bytes > workspace_bytes) return VB_STATUS_BAD_ARGUMENT;
    double* hs = pinned_status();
    VB_CHECK(cudaMemsetAsync(w.sc, 0, CG_NSCAL * sizeof(double), st));
    {   // weighted degrees = diagonal of J^T J (also the Jacobi preconditioner of the accurate mode)
        if (n_t > 0) seg_sum1_kernel<<<tr_warp_grid(n_t), TR_THREADS, 0, st>>>(g->t_rowptr, nullptr, g->t_w, w.dg_t, n_t);
        cam_runs_sum_kernel<<<tr_warp_grid(n_c), TR_THREADS, 0, st>>>(g->c_segptr, g->n_windows, n_c, nullptr, g->c_w, w.dg_c);
        if (allreduce) { int rc = allreduce(actx, w.dg_c, n_c, (void*)st); if (rc) return rc; }
    }
    // camera part is replicated across ranks -> counted once on every rank; time part is local
    cg_init_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(rhs_c, w.dg_c, jacobi, x_c, w.r_c, w.p_c, n_c, w.sc, CG_RHO_NEXT_C, CG_RN2_C);
    if (n_t > 0) cg_init_kernel<<<tr_grid(n_t), TR_THREADS, 0, st>>>(rhs_t, w.dg_t, jacobi, x_t, w.r_t, w.p_t, n_t, w.sc, CG_RHO_NEXT_T, CG_RN2_T);
    VB_KERNEL_CHECK();
    double* pack = w.q_c + 3 * n_c;
    auto reduce_scalars = [&](int s0, int s1, int s2) -> int {
        if (!allreduce) return 0;
        pack_scalars_kernel<<<1, 1, 0, st>>>(pack, w.sc, s0, s1, s2);
        int rc = allreduce(actx, pack, 3, (void*)st);
        if (rc) return rc;
        unpack_scalars_kernel<<<1, 1, 0, st>>>(pack, w.sc, s0, s1, s2);
        return 0;
    };
    { int rc = reduce_scalars(CG_RHO_NEXT_T, CG_RN2_T, -1); if (rc) return rc; }
    int status = VB_STATUS_NOT_CONVERGED;
    const int check_every = 4;
    for (int64_t it = 0; it <= maxiter; ++it) {
        cg_scalar_top_kernel<<<1, 1, 0, st>>>(w.sc, rtol, it == 0 ? 1 : 0);
        if (it % check_every == 0 || it == maxiter) {
            VB_CHECK(cudaMemcpyAsync(hs, w.sc, CG_NSCAL * sizeof(double), cudaMemcpyDeviceToHost, st));
            VB_CHECK(cudaStreamSynchronize(st));
            if (hs[CG_DONE] != 0.0) { status = VB_STATUS_OK; break; }
        }
        if (it == maxiter) break;
        cg_dir_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(w.r_c, w.dg_c, jacobi, w.p_c, n_c, w.sc);
        if (n_t > 0) cg_dir_kernel<<<tr_grid(n_t), TR_THREADS, 0, st>>>(w.r_t, w.dg_t, jacobi, w.p_t, n_t, w.sc);
        cg_qc_init_kernel<<<tr_grid(n_c < 8 ? 8 : n_c), TR_THREADS, 0, st>>>(w.dg_c, w.p_c, w.q_c, n_c, owner, w.sc);
        if (n_t > 0) {
            int tg = tr_warp_grid(n_t);
            const int cap = sm_count() * 8;   // persistent: 8 CTAs of 256 threads per SM
            if (tg > cap) tg = cap;
            cg_time_kernel<<<tg, TR_THREADS, 0, st>>>(g->t_rowptr, g->t_cam, g->t_w, w.dg_t, w.p_c, w.p_t, w.q_t, n_t, w.sc);
        }
        if (g->n_tiles > 0) cg_cam_kernel<<<tr_warp_grid(g->n_tiles), TR_THREADS, 0, st>>>(g->tile_cam, g->tile_start, g->tile_end, g->c_time, g->c_w, w.p_c, w.p_t, w.q_c, g->n_tiles, w.sc);
        VB_KERNEL_CHECK();
        if (allreduce) {
            // one collective: camera accumulator + the local time part of p.q packed behind it
            pack_scalars_kernel<<<1, 1, 0, st>>>(pack, w.sc, CG_PQ_T, -1, -1);
            int rc = allreduce(actx, w.q_c, 3 * n_c + 8, (void*)st);
            if (rc) return rc;
            unpack_scalars_kernel<<<1, 1, 0, st>>>(pack, w.sc, CG_PQ_T, -1, -1);
        }
        cg_dot_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(w.p_c, w.q_c, n_c, w.sc, CG_PQ_C);
        cg_scalar_alpha_kernel<<<1, 1, 0, st>>>(w.sc);
        cg_update_kernel<<<tr_grid(n_c), TR_THREADS, 0, st>>>(w.p_c, w.q_c, w.dg_c, jacobi, x_c, w.r_c, n_c, w.sc, CG_RHO_NEXT_C, CG_RN2_C);
        if (n_t > 0) cg_update_kernel<<<tr_grid(n_t), TR_THREADS, 0, st>>>(w.p_t, w.q_t, w.dg_t, jacobi, x_t, w.r_t, n_t, w.sc, CG_RHO_NEXT_T, CG_RN2_T);
        VB_KERNEL_CHECK();
        { int rc = reduce_scalars(CG_RHO_NEXT_T, CG_RN2_T, -1); if (rc) return rc; }
    }
    if (h_iters) *h_iters = (int32_t)hs[CG_ITERS];
    return status;
}

}  // namespace vb
