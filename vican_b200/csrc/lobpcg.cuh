// Width-3 LOBPCG step for the 3 eigenpairs of L = Lambda_C - P Lambda_T P^T nearest zero.
//
// Replaces scipy.sparse.linalg.eigs(L, k=5, sigma=-1e-6) (ARPACK shift-invert + SuperLU,
// vican/bipgo.py:288): only the 3-dimensional invariant subspace enters the result
// (bipgo.py:295), and it is independent of the eigensolver.
//
// One cooperative kernel does ALL vector work between two applications of L:
//   stage 1  A-image of the freshly applied block (AW = Lambda_C W - Y), Gram matrices
//            G = S^T A S and M = S^T S of the basis S = [X W P]                  (grid reduce 1)
//   stage 2  9x9 Rayleigh-Ritz, redundantly per CTA (dense_small.cuh: ritz9)
//   stage 3  X,AX,P,AP <- basis update; residual R = AX - X theta; W = Lambda_C^-1 R
//            (block-Jacobi preconditioner); H = [X P]^T W, ||R_j||^2           (grid reduce 2)
//            -> convergence test (uniform across CTAs)
//   stage 4  W -= [X P] H; second projection coefficients and Gram of W         (grid reduce 3)
//   stage 5  W -= [X P] H2; W <- W T  (SVQB orthonormalisation with dropping)
// Block vectors are stored as [n_c][9] (camera block, row-major 3x3): row (c,i), column j.
// Camera-side vectors are tiny (72 B per camera per vector), so this kernel is latency- not
// bandwidth-bound; it exists to keep the whole eigen-iteration on the device.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "dense_small.cuh"

namespace vb {
namespace cg = cooperative_groups;

constexpr int LOB_THREADS = 256;
constexpr int LOB_NRED = 112;      // >= 12*9 + 1
// persistent small state (doubles)
constexpr int SM_THETA = 0, SM_RESN = 3, SM_CONV = 6, SM_ANORM = 7, SM_ITERS = 8, SM_ACT = 9, SM_TIME = 18, SM_NCONV = 29, SM_SIZE = 32;
// SM_NCONV = 1 - SM_CONV: skip flag of launches that are only valid once the step HAS converged (so3sync_run)
// SM_TIME..+9: globaltimer stamps (ns) of block 0 at the stage boundaries of the last step (diagnostics)

struct LobpcgParams {
    int n_c;
    double *X, *AX, *W, *AW, *P, *AP;
    const double* Y;
    const double *lamC, *lamCinv;
    double* small;
    double* partial;   // [3][gridDim.x][LOB_NRED]
    double* Wpad;      // new W also in the padded gather layout [n_c][GSTRIDE] (input of the next time pass)
    double tol;
    int first;
};

// ---- warp reduce-scatter: sums N per-lane values over the 32 lanes with a halving butterfly.
// At the step with lane mask MASK every lane keeps one half of its live values and ships the other
// half to its partner, so the whole reduction costs about N fp64 shuffles instead of 5 N (shuffles
// share the LSU pipe with shared/global loads; a plain all-reduce of the 108 Gram entries was
// the longest part of the vector stages).  After the last step a lane holds the totals of the values
// [base, base + count) of the original array (entries beyond N are padding) and emits them.
template <int N, int MASK>
struct RsStep {
    static constexpr int H = (N + 1) / 2;
    // `valid` = how many of the N entries are real (the upper half of an odd split carries one
    // padding slot whose index belongs to the NEXT region: it must never be emitted)
    template <typename F>
    __device__ __forceinline__ static void run(const double (&v)[N], int lane, int base, int valid, F&& emit) {
        const bool up = (lane & MASK) != 0;
        double w[H];
#pragma unroll
        for (int m = 0; m < H; ++m) {
            const double lo = v[m];
            const double hi = (H + m < N) ? v[H + m] : 0.0;
            const double recv = shfl_xor(up ? lo : hi, MASK);
            w[m] = (up ? hi : lo) + recv;
        }
        const int nbase = base + (up ? H : 0);
        const int nvalid = up ? (valid > H ? valid - H : 0) : (valid < H ? valid : H);
        if constexpr (MASK > 1) RsStep<H, MASK / 2>::run(w, lane, nbase, nvalid, emit);
        else {
#pragma unroll
            for (int m = 0; m < H; ++m)
                if (m < nvalid) emit(nbase + m, w[m]);
        }
    }
};

// block-level sum of N per-thread values -> dst[0..N) (one slot per CTA, deterministic order):
// reduce-scatter inside every warp, holders drop their totals into shared memory, N threads add the
// per-warp rows.  `sm` must hold (LOB_THREADS / 32) * N doubles.  Ends with a __syncthreads().
template <int N>
__device__ __forceinline__ void block_sum_store(const double (&v)[N], double* sm, double* dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* row = sm + warp * N;
    RsStep<N, 16>::run(v, lane, 0, N, [&](int idx, double tot) { row[idx] = tot; });
    __syncthreads();
    if (threadIdx.x < N) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < LOB_THREADS / 32; ++w) t += sm[w * N + threadIdx.x];
        dst[threadIdx.x] = t;
    }
    __syncthreads();
}

// deterministic: every CTA sums the per-CTA partials in the same order.  Lanes run along the value
// index (coalesced rows of the partial table; a lane-per-CTA layout reads one 32-byte sector per
// double and was 4x slower), the 8 warps split the CTAs, shared memory joins the 8 strands.
__device__ __forceinline__ void grid_combine(const double* partial, int n, double* out_sm, double* comb_sm /*[8][LOB_NRED]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = LOB_THREADS / 32;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = warp; b < (int)gridDim.x; b += NW) {
        const double* row = partial + (size_t)b * LOB_NRED;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (lane + 32 * u < n) s[u] += row[lane + 32 * u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (lane + 32 * u < n) comb_sm[warp * LOB_NRED + lane + 32 * u] = s[u];
    __syncthreads();
    if ((int)threadIdx.x < n) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += comb_sm[w * LOB_NRED + threadIdx.x];
        out_sm[threadIdx.x] = t;
    }
    __syncthreads();
}

__device__ __forceinline__ void ld3(const double* p, double* v) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
__device__ __forceinline__ void st3(double* p, const double* v) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; }
// out[j] += sum_jj v[jj] * Cm[jj * 3 + j]     (row vector times 3x3)
__device__ __forceinline__ void row_times(const double* v, const double* Cm, double* out) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] += v[0] * Cm[j] + v[1] * Cm[3 + j] + v[2] * Cm[6 + j];
}
// acc[3j+jp] += a[j] * b[jp]      (contribution of one row to a^T b)
__device__ __forceinline__ void outer_acc(const double* a, const double* b, double* acc) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int jp = 0; jp < 3; ++jp) acc[3 * j + jp] += a[j] * b[jp];
}
// row i of (block-diagonal 3x3) * (3 rows of the camera): out[j] = sum_k l[k] * V[3 k + j]
__device__ __forceinline__ void blockrow_times(const double* l, const double* V9, double* out) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] = l[0] * V9[j] + l[1] * V9[3 + j] + l[2] * V9[6 + j];
}

__device__ __forceinline__ void lob_stamp(const LobpcgParams& p, int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.small[SM_TIME + k] = (double)t;
    }
}

// Work decomposition: one thread per ROW of the block vectors (row r = 3 c + i of camera c; a
// block vector [n_c][9] is a row-major [3 n_c][3] matrix, so consecutive threads read consecutive
// 24-byte rows: coalesced, three times the parallelism of a thread per camera).  A CTA owns
// LOB_ROWS = 255 rows per chunk = 85 whole cameras, so the two block-diagonal products
// (Lambda_C V and Lambda_C^-1 R read the three rows of a camera) never cross a CTA.
constexpr int LOB_ROWS = 255;

__global__ void __launch_bounds__(LOB_THREADS, 1) lobpcg_step_kernel(LobpcgParams p) {
    // a speculatively enqueued step after convergence is a no-op (uniform across the grid: the flag
    // was written by the previous launch)
    if (!p.first && p.small[SM_CONV] != 0.0) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ double red_sm[(LOB_THREADS / 32) * 27];
    __shared__ double tot[LOB_NRED];
    __shared__ double comb_sm[(LOB_THREADS / 32) * LOB_NRED];
    __shared__ double Gm[81], Mm[81], Cx[27], Cp[27], work[5 * 81 + 64];
    __shared__ int iwork[48];
    __shared__ double theta_s[3], H[18], T[9];
    __shared__ int act_s[9], actP_s[3], actW_s[3], conv_s;

    const int n_rows = 3 * p.n_c;
    const int row0 = blockIdx.x * LOB_ROWS + threadIdx.x;          // first row of this thread
    const int rstride = gridDim.x * LOB_ROWS;
    const bool rower = threadIdx.x < LOB_ROWS;                     // thread 255 only helps in the reductions
    const int irow = threadIdx.x % 3;                              // row inside the camera block (LOB_ROWS % 3 == 0)
    const double* base1 = p.partial;
    const double* base2 = p.partial + (size_t)gridDim.x * LOB_NRED;
    const double* base3 = p.partial + (size_t)2 * gridDim.x * LOB_NRED;
    double* part1 = p.partial + (size_t)blockIdx.x * LOB_NRED;
    double* part2 = p.partial + ((size_t)gridDim.x + blockIdx.x) * LOB_NRED;
    double* part3 = p.partial + ((size_t)2 * gridDim.x + blockIdx.x) * LOB_NRED;

    lob_stamp(p, 0);
    // ---------------- stage 1: A-image of the applied block, Gram matrices -----------------
    {
        double* V = p.first ? p.X : p.W;
        double* AV = p.first ? p.AX : p.AW;
        double an = 0.0;
        if (rower)
            for (int r = row0; r < n_rows; r += rstride) {
                double v[9], y[3], l[3], o[3];
                const double* vc = V + 3 * (size_t)(r - irow);     // the camera's three rows
#pragma unroll
                for (int q = 0; q < 9; ++q) v[q] = vc[q];
                ld3(p.Y + 3 * (size_t)r, y);
                ld3(p.lamC + 3 * (size_t)r, l);
                blockrow_times(l, v, o);
#pragma unroll
                for (int j = 0; j < 3; ++j) o[j] -= y[j];
                an += l[0] * l[0] + l[1] * l[1] + l[2] * l[2];
                st3(AV + 3 * (size_t)r, o);
            }
        // the row stored above is re-read below by the same thread only -> no sync needed
        const double* Vs[3] = {p.X, p.W, p.P};
        const double* AVs[3] = {p.AX, p.AW, p.AP};
        int blk = 0;
        for (int a = 0; a < 3; ++a)
            for (int b = a; b < 3; ++b, ++blk) {
                double gm[19];
#pragma unroll
                for (int q = 0; q < 19; ++q) gm[q] = 0.0;
                if (rower)
                    for (int r = row0; r < n_rows; r += rstride) {
                        double va[3], vb_[3], avb[3];
                        ld3(Vs[a] + 3 * (size_t)r, va);
                        ld3(Vs[b] + 3 * (size_t)r, vb_);
                        ld3(AVs[b] + 3 * (size_t)r, avb);
                        outer_acc(va, avb, gm);
                        outer_acc(va, vb_, gm + 9);
                    }
                gm[18] = (blk == 0) ? an : 0.0;
                // 19 totals of this block pair: G block -> part1[9 blk ..], M block -> part1[54 + 9 blk ..], |Lambda_C|^2 -> [108]
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                double* row = red_sm + warp * 19;
                RsStep<19, 16>::run(gm, lane, 0, 19, [&](int idx, double t) { row[idx] = t; });
                __syncthreads();
                if (threadIdx.x < 19) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < LOB_THREADS / 32; ++w) t += red_sm[w * 19 + threadIdx.x];
                    const int q = threadIdx.x;
                    if (q < 9) part1[9 * blk + q] = t;
                    else if (q < 18) part1[54 + 9 * blk + (q - 9)] = t;
                    else if (blk == 0) part1[108] = t;
                }
                __syncthreads();
            }
    }
    lob_stamp(p, 1);
    grid.sync();
    lob_stamp(p, 2);
    grid_combine(base1, 109, tot, comb_sm);
    lob_stamp(p, 3);

    // ---------------- stage 2: Rayleigh-Ritz (redundant per CTA, deterministic) ------------
    // one warp per CTA: warp-cooperative 9x9 solve (parallel-ordered Jacobi), dense_small.cuh
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int idx = lane; idx < 54; idx += 32) {
            const int blk = idx / 9, e9 = idx - 9 * blk, j = e9 / 3, jp = e9 - 3 * j;
            // upper blocks in the order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
            const int a = (blk < 3) ? 0 : ((blk < 5) ? 1 : 2);
            const int b = (blk < 3) ? blk : ((blk < 5) ? blk - 2 : 2);
            const double g = tot[idx], m = tot[54 + idx];
            Gm[(3 * a + j) * 9 + 3 * b + jp] = g;
            Mm[(3 * a + j) * 9 + 3 * b + jp] = m;
            if (a != b) {
                Gm[(3 * b + jp) * 9 + 3 * a + j] = g;
                Mm[(3 * b + jp) * 9 + 3 * a + j] = m;
            }
        }
        for (int i = lane; i < 9; i += 32) act_s[i] = (i < 3) ? 1 : (p.first ? 0 : (p.small[SM_ACT + i] != 0.0));
        __syncwarp();
        ritz9_coop(Gm, Mm, act_s, Cx, Cp, theta_s, actP_s, work, iwork, lane, 32);
        if (lane == 0 && blockIdx.x == 0) p.small[SM_TIME + 10] = (double)iwork[30];   // Jacobi sweeps (diagnostics)
    }
    __syncthreads();
    lob_stamp(p, 4);

    // ---------------- stage 3: basis update, residual, preconditioned direction ------------
    const double anorm = sqrt(tot[108] / (3.0 * p.n_c));
    {
        // 3a (row-local, in place): X, P, AX, AP <- basis update; the residual row R = AX - X theta is parked in
        // the AW array (A W is dead from here on; the next step recomputes it)
        if (rower)
            for (int r = row0; r < n_rows; r += rstride) {
                const size_t o = 3 * (size_t)r;
                double x[3], w[3], pp[3], xn[3] = {0, 0, 0}, pn[3] = {0, 0, 0};
                ld3(p.X + o, x); ld3(p.W + o, w); ld3(p.P + o, pp);
                row_times(x, Cx, xn);  row_times(w, Cx + 9, xn);  row_times(pp, Cx + 18, xn);
                row_times(x, Cp, pn);  row_times(w, Cp + 9, pn);  row_times(pp, Cp + 18, pn);
                st3(p.X + o, xn); st3(p.P + o, pn);
                double ax[3], aw[3], ap[3], axn[3] = {0, 0, 0}, apn[3] = {0, 0, 0}, res[3];
                ld3(p.AX + o, ax); ld3(p.AW + o, aw); ld3(p.AP + o, ap);
                row_times(ax, Cx, axn);  row_times(aw, Cx + 9, axn);  row_times(ap, Cx + 18, axn);
                row_times(ax, Cp, apn);  row_times(aw, Cp + 9, apn);  row_times(ap, Cp + 18, apn);
                st3(p.AX + o, axn); st3(p.AP + o, apn);
#pragma unroll
                for (int j = 0; j < 3; ++j) res[j] = axn[j] - xn[j] * theta_s[j];
                st3(p.AW + o, res);
            }
        __syncthreads();   // the three rows of a camera live in one CTA
        // 3b: W = Lambda_C^-1 R (block-Jacobi preconditioner), H = [X P]^T W, ||R_j||^2
        double acc[21];
#pragma unroll
        for (int q = 0; q < 21; ++q) acc[q] = 0.0;
        if (rower)
            for (int r = row0; r < n_rows; r += rstride) {
                const size_t o = 3 * (size_t)r;
                double R9[9], li[3], wn[3], xn[3], pn[3], rr[3];
                const double* rc = p.AW + 3 * (size_t)(r - irow);
                ld3(p.AW + o, rr);                                   // own residual row (no dynamic register indexing)
#pragma unroll
                for (int q = 0; q < 9; ++q) R9[q] = rc[q];
                ld3(p.lamCinv + o, li);
                blockrow_times(li, R9, wn);
                st3(p.W + o, wn);
                ld3(p.X + o, xn); ld3(p.P + o, pn);
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[18 + j] += rr[j] * rr[j];
                outer_acc(xn, wn, acc);
                outer_acc(pn, wn, acc + 9);
            }
        block_sum_store<21>(acc, red_sm, part2);
    }
    lob_stamp(p, 5);
    grid.sync();
    grid_combine(base2, 21, tot, comb_sm);
    lob_stamp(p, 6);
    if (threadIdx.x == 0) {
        const double r0 = sqrt(tot[18]), r1 = sqrt(tot[19]), r2 = sqrt(tot[20]);
        const double rmax = fmax(r0, fmax(r1, r2));
        conv_s = (rmax <= p.tol * anorm) ? 1 : 0;
        for (int i = 0; i < 18; ++i) H[i] = tot[i];
        if (blockIdx.x == 0) {
            p.small[SM_THETA] = theta_s[0]; p.small[SM_THETA + 1] = theta_s[1]; p.small[SM_THETA + 2] = theta_s[2];
            p.small[SM_RESN] = r0; p.small[SM_RESN + 1] = r1; p.small[SM_RESN + 2] = r2;
            p.small[SM_CONV] = (double)conv_s;
            p.small[SM_NCONV] = (double)(1 - conv_s);
            p.small[SM_ANORM] = anorm;
            p.small[SM_ITERS] = p.first ? 1.0 : p.small[SM_ITERS] + 1.0;
            for (int j = 0; j < 3; ++j) p.small[SM_ACT + 6 + j] = (double)actP_s[j];
        }
    }
    __syncthreads();
    if (conv_s) return;   // uniform over the grid: every CTA evaluated the same numbers

    // ---------------- stage 4: project W against [X P], Gram of W ---------------------------
    {
        double acc[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) acc[q] = 0.0;
        if (rower)
            for (int r = row0; r < n_rows; r += rstride) {
                const size_t o = 3 * (size_t)r;
                double x[3], pp[3], w[3];
                ld3(p.X + o, x); ld3(p.P + o, pp); ld3(p.W + o, w);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double sum = 0.0;
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) sum += x[jj] * H[3 * jj + j] + pp[jj] * H[9 + 3 * jj + j];
                    w[j] -= sum;
                }
                st3(p.W + o, w);
                outer_acc(x, w, acc);
                outer_acc(pp, w, acc + 9);
                outer_acc(w, w, acc + 18);
            }
        block_sum_store<27>(acc, red_sm, part3);
    }
    lob_stamp(p, 7);
    grid.sync();
    grid_combine(base3, 27, tot, comb_sm);
    lob_stamp(p, 8);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 18; ++i) H[i] = tot[i];
        // Gram of the twice-projected block: the second correction is O(eps) so
        // W2^T W2 = Gw - H2^T H2 to O(eps^2); keep it symmetric.
        double Gw[9];
        for (int j = 0; j < 3; ++j)
            for (int jp = 0; jp < 3; ++jp) {
                double sum = tot[18 + 3 * j + jp];
                for (int k = 0; k < 6; ++k) sum -= tot[3 * k + j] * tot[3 * k + jp];
                Gw[3 * j + jp] = sum;
            }
        svqb3(Gw, T, actW_s, 1e-12);
        if (blockIdx.x == 0)
            for (int j = 0; j < 3; ++j) p.small[SM_ACT + 3 + j] = (double)actW_s[j];
    }
    __syncthreads();

    // ---------------- stage 5: second projection + orthonormalisation ------------------------
    if (rower)
        for (int r = row0; r < n_rows; r += rstride) {
            const size_t o = 3 * (size_t)r;
            double x[3], pp[3], w[3], wn[3] = {0, 0, 0};
            ld3(p.X + o, x); ld3(p.P + o, pp); ld3(p.W + o, w);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double sum = 0.0;
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) sum += x[jj] * H[3 * jj + j] + pp[jj] * H[9 + 3 * jj + j];
                w[j] -= sum;
            }
            row_times(w, T, wn);
            st3(p.W + o, wn);
            if (p.Wpad) { double* q = p.Wpad + GSTRIDE * (size_t)(r / 3) + 4 * irow; q[0] = wn[0]; q[1] = wn[1]; q[2] = wn[2]; q[3] = 0.0; }
        }
    lob_stamp(p, 9);
}

inline int lobpcg_grid(int n_c) {
    int want = (3 * n_c + LOB_ROWS - 1) / LOB_ROWS;   // one thread per row, 255 rows (85 cameras) per CTA
    const int cap = sm_count();   // cooperative launch: all CTAs co-resident (1 CTA / SM)
    if (want < 1) want = 1;
    return want < cap ? want : cap;
}

inline int launch_lobpcg_step(const LobpcgParams& p, cudaStream_t st) {
    LobpcgParams pp = p;
    void* args[] = {(void*)&pp};
    VB_CHECK(cudaLaunchCooperativeKernel((void*)lobpcg_step_kernel, dim3(lobpcg_grid(p.n_c)), dim3(LOB_THREADS), args, 0, st));
    return 0;
}

}  // namespace vb
