// Edge passes of the matrix-free SE(3) Laplacian (the HBM-bound hot kernels).
//
// The reference forms the 3n_c x 3n_c power-graph matrix P Lambda_T P^T with an SpGEMM every
// outer iteration and multiplies by it with scipy's CSR kernels (vican/bipgo.py:273, :300,
// :318, :334).  Here nothing is formed.  One application of  P Lambda_T P^T  to a block
// vector X (n_c blocks of 3x3) is two streaming passes over the aggregated edge blocks:
//
//   time pass  (edges sorted by time node):  Z_t = sum_{e in t} B_e^T X_{c_e},  W_t = Lambda_T[t] Z_t
//   camera pass(edges in (window, camera) tiles): Y_c += sum_{e in tile} B_e W_{t_e}
//
// Both are *gathers* on the far endpoint and segmented reductions on the owning endpoint, so
// no fp64 scatter-atomics per edge are needed (450 M atomics per pass at 50 M edges would be
// 3.5x slower than streaming the blocks; shared-memory fp64 atomics are CAS loops).
//
// Kernel structure (third generation; profiles/r1_edge_pass_history.md has the measurements
// that led here):
//  * every warp is an independent pipeline that owns segments warp, warp+W, ... (a time node or
//    a camera tile).  One elected lane streams the NEXT work item (<= 50 edges: their 72-byte
//    blocks and 4-byte indices) into the warp's double-buffered shared-memory stage with two
//    cp.async.bulk copies (TMA, mbarrier completion, L2 evict-first) while the warp consumes
//    the current item.  No CTA-wide synchronisation exists.
//  * lane mapping: 30 lanes = 10 edges x 3 rows.  Lane (e, k) reads row k of edge e's staged
//    block from shared memory (3 LDS.64, conflict-free; the camera-pass copy of the blocks is
//    stored transposed, so its "row k" is column k of B_e) and row k of the gathered node block
//    with ONE 256-bit load (node blocks are stored padded: 3 rows of 4 doubles in one 128-byte
//    line, so a row is 32-byte aligned and an edge touches one L1 line), then does the 9 FMAs of
//    B[k][:]^T x G[k][:] into a private 3x3 accumulator.  No shuffles in the inner loop: the
//    previous mapping (9 lanes per edge, row broadcast by shuffles) was bound by the LSU /
//    shuffle return path at 52-60 % of the HBM roofline.
//  * per segment the 30 private 3x3 sums are combined with a reduce-scatter butterfly
//    (12 fp64 shuffles instead of 45 for a plain all-reduce).
//
// Algorithmic bytes per edge: 72 (block) + 4 (index) = 76 B  (SURVEY.md 8d).
#pragma once
#include <stdlib.h>

#include "../../include/vican_b200.h"
#include "common.cuh"
#include "peer.cuh"

namespace vb {


constexpr int PASS_THREADS = 128;          // 4 independent warp pipelines per CTA
constexpr int PASS_CTAS_PER_SM = 4;        // 4 x ~42 KB shared memory, <= 127 registers
constexpr int PASS_WARPS = PASS_THREADS / 32;
constexpr int ITEM_EDGES = 50;                                        // one work item = 5 rounds of 10 edges
constexpr int BUF_B_BYTES = (ITEM_EDGES + 2) * 72;                    // 3744, multiple of 16
constexpr int BUF_I_BYTES = 240;                                      // >= (ITEM_EDGES + 6) * 4, multiple of 16
constexpr int BUF_BYTES = ((BUF_B_BYTES + BUF_I_BYTES + 127) / 128) * 128;   // 4096
constexpr int WARP_SCRATCH = 128;                                     // 9 doubles of epilogue scratch
constexpr int WARP_SMEM = 2 * BUF_BYTES + WARP_SCRATCH;
constexpr int PASS_SMEM = PASS_WARPS * WARP_SMEM + PASS_WARPS * 2 * 8;
constexpr int EDGES_PER_ROUND = 10;
constexpr int PASS_VAR_DEFAULT = 0;

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
// one 32-byte row of a padded node block; gathered blocks are re-read by many edges -> keep in L2
// (the default L2 policy instead of evict-last measured 3-5 % slower inside the solve)
__device__ __forceinline__ void ld_row256(const double* p, double& a, double& b, double& c) {
    double pad;   // 4th double of the row: padding
    asm volatile("ld.global.nc.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(pad) : "l"(p));
    (void)pad;
}

// Sum 9 values over the 32 lanes with a reduce-scatter butterfly: at every step a lane keeps
// (roughly) half of its live values and ships the other half to its partner -> 5+3+2+1+1 = 12
// fp64 shuffles.  Returns the lane's total; *vidx = which of the 9 values it is (0xF = none).
// Value m lives in both lanes of one pair (see table); the even lane acts as its holder.
__device__ __forceinline__ double reduce_scatter9(const double (&v)[9], int lane, int* vidx) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    double w[5], u[3], t[2], s;
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const double hi = (m < 4) ? v[5 + m] : 0.0;
        const double recv = shfl_xor(b4 ? v[m] : hi, 16);
        w[m] = (b4 ? hi : v[m]) + recv;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const double hi = (m < 2) ? w[3 + m] : 0.0;
        const double recv = shfl_xor(b3 ? w[m] : hi, 8);
        u[m] = (b3 ? hi : w[m]) + recv;
    }
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const double hi = (m < 1) ? u[2 + m] : 0.0;
        const double recv = shfl_xor(b2 ? u[m] : hi, 4);
        t[m] = (b2 ? hi : u[m]) + recv;
    }
    {
        const double recv = shfl_xor(b1 ? t[0] : t[1], 2);
        s = (b1 ? t[1] : t[0]) + recv;
    }
    s += shfl_xor(s, 1);
    // (b4 b3 b2 b1) -> value index; 0xF marks lanes whose slot is padding of the odd splits
    const unsigned long long tbl = 0xFFF8F765FF43F210ULL;
    *vidx = (int)((tbl >> (4 * ((lane >> 1) & 15))) & 0xF);
    return s;
}

struct Item {
    int seg, a, b;
    bool last, valid;
};

// item loads: block rows (LDS) + far-endpoint indices (LDS) + 256-bit row gathers, all issued back to back.
// Lanes without an edge in a round get a ZERO block row (their gathered row keeps whatever finite value the
// registers held: x is cleared once before the loop), so the FMAs of a round need only a warp-uniform guard.
__device__ __forceinline__ void item_load(const unsigned char* bufp, const double* __restrict__ G, const Item& it, int e,
                                          int k, bool lane_on, double (&b)[5][3], double (&x)[5][3]) {
    const double* sB = reinterpret_cast<const double*>(bufp);
    const int* sI = reinterpret_cast<const int*>(bufp + BUF_B_BYTES);
    const int offB = it.a - (it.a & ~1), offI = it.a - (it.a & ~3), n_e = it.b - it.a;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
        const int off = EDGES_PER_ROUND * u + e;
        b[u][0] = b[u][1] = b[u][2] = 0.0;
        if (lane_on && (off < n_e)) {
            const int node = sI[offI + off];
            // row k of the staged block: row k of B for the time pass, row k of B^T (= column k of B) for the
            // camera pass, whose copy of the blocks is stored transposed
            const double* pb = sB + 9 * (offB + off) + 3 * k;
            b[u][0] = pb[0];
            b[u][1] = pb[1];
            b[u][2] = pb[2];
            ld_row256(G + GSTRIDE * (size_t)node + 4 * k, x[u][0], x[u][1], x[u][2]);
        }
    }
}

__device__ __forceinline__ void item_fma(const Item& it, const double (&b)[5][3], const double (&x)[5][3], double (&acc)[9]) {
    const int n_e = it.b - it.a;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
        if (EDGES_PER_ROUND * u < n_e) {   // warp-uniform: skip empty tail rounds
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[3 * a + j] = fma(b[u][a], x[u][j], acc[3 * a + j]);
        }
    }
}

// MODE 0: out_t = Lambda_T[t] * sum B^T X   (padded rows)   -- L-apply / primal multiply (bipgo.py:300)
// (raw gather out_t = sum B^T X for the dual update, bipgo.py:318: MODE 0 with lamT == nullptr, i.e. Lambda_T = I)
// MODE 2: Y_c  += sum over tile of B W      (compact 9, fp64 atomics per TILE, not per edge: the fused multi-GPU pass)
// MODE 3: part[tile] = sum over tile of B W (plain stores; tile_combine_kernel adds a camera's tiles in a fixed
//         order: the single-GPU camera pass is bitwise reproducible and needs no zeroed accumulator)
//
// Per-warp software pipeline over work items (<= 50 edges of one segment):
//   FMA(item k)  ->  TMA issue(item k+2)  ->  wait + loads(item k+1)  ->  epilogue(segment of k, if it ends)
// so the segment reduction / store overlaps the gather round trip of the next item and the TMA copy
// of an item has one full iteration to land.
// VAR (tuning variants, selected per launch): bit 0 = quad-aligned lane mapping (else dense lane = 3 e + k),
// bit 1 = MODE 0 epilogue through shuffles (else through the per-warp scratch in shared memory).
template <int MODE, int VAR>
__device__ __forceinline__ void edge_pass_body(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node,
                                               const int* __restrict__ idx, const double* __restrict__ B,
                                               const double* __restrict__ G, const double* __restrict__ lamT,
                                               double* __restrict__ out, int n_seg) {
    extern __shared__ __align__(128) unsigned char pass_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp = blockIdx.x * PASS_WARPS + wib;
    const int nwarps = gridDim.x * PASS_WARPS;
    unsigned char* wbuf = pass_smem + (size_t)wib * WARP_SMEM;
    double* scratch = reinterpret_cast<double*>(wbuf + 2 * BUF_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(pass_smem + (size_t)PASS_WARPS * WARP_SMEM) + 2 * wib;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (warp >= n_seg) return;

    const uint64_t pf = policy_evict_first();
    // Lane -> (edge of the round, block row).  Quad q (lanes 4q .. 4q+3) holds the three rows of edge q in its
    // lanes 0-2: their 256-bit gathers hit ONE 128-byte line (a padded node block), i.e. one L1 wavefront.
    // The fourth lanes of the quads carry the six rows of edges 8 and 9 (two stay idle), assigned so that the
    // LDS.64 of the staged block rows stay bank-conflict free.  10 edges per round at 14 gather wavefronts
    // (a dense lane = 3 e + k packing straddles quads: 15-19 wavefronts per round, measured 0.63 per row).
    const int quad = lane >> 2, qr = lane & 3;
    int e, k;
    bool lane_on = true;
    if (!(VAR & 1)) { e = lane / 3; k = lane - 3 * e; lane_on = lane < 30; }
    else if (qr < 3) { e = quad; k = qr; }
    else {
        // quad:      0      1      2     3     4      5      6      7
        // carries: (9,1)  (9,2)   idle  idle  (8,0)  (8,1)  (8,2)  (9,0)
        e = (int)((0x98880099u >> (4 * quad)) & 0xFu);
        k = (int)((0x02100021u >> (4 * quad)) & 0xFu);
        lane_on = (quad != 2) && (quad != 3);
    }

    // ---- item iterator (warp-uniform).  Row pointers are fetched two segments ahead of their use.
    int it_seg = warp, it_s = __ldg(seg_ptr + warp), it_e = __ldg(seg_ptr + warp + 1), it_pos = it_s;
    bool it_started = false, it_valid = true;
    int n1_seg = warp + nwarps, n1_s = 0, n1_e = 0, n2_s = 0, n2_e = 0;
    if (n1_seg < n_seg) { n1_s = __ldg(seg_ptr + n1_seg); n1_e = __ldg(seg_ptr + n1_seg + 1); }
    if (n1_seg + nwarps < n_seg) { n2_s = __ldg(seg_ptr + n1_seg + nwarps); n2_e = __ldg(seg_ptr + n1_seg + nwarps + 1); }
    auto next_item = [&]() -> Item {
        Item it;
        it.seg = 0; it.a = 0; it.b = 0; it.last = false; it.valid = false;
        if (!it_valid) return it;
        if (it_started && it_pos >= it_e) {
            if (n1_seg >= n_seg) { it_valid = false; return it; }
            it_seg = n1_seg; it_s = n1_s; it_e = n1_e; it_pos = it_s; it_started = false;
            n1_seg += nwarps; n1_s = n2_s; n1_e = n2_e;
            if (n1_seg + nwarps < n_seg) { n2_s = __ldg(seg_ptr + n1_seg + nwarps); n2_e = __ldg(seg_ptr + n1_seg + nwarps + 1); }
        }
        it.seg = it_seg; it.a = it_pos; it.b = min(it_pos + ITEM_EDGES, it_e);
        it_pos = it.b; it_started = true;
        it.last = (it.b >= it_e); it.valid = true;
        return it;
    };
    auto issue = [&](const Item& it, int buf) {   // whole warp calls; one elected lane issues
        if (!it.valid || it.b <= it.a) return;
        __syncwarp();
        if (lane == 0) {
            const int a0 = it.a & ~1, b1 = (it.b + 1) & ~1;
            const int a0i = it.a & ~3, b1i = (it.b + 3) & ~3;
            const uint32_t nbB = (uint32_t)(b1 - a0) * 72u, nbI = (uint32_t)(b1i - a0i) * 4u;
            unsigned char* dst = wbuf + buf * BUF_BYTES;
            mbar_expect_tx(&bars[buf], nbB + nbI);
            bulk_g2s(dst, B + 9 * (size_t)a0, nbB, &bars[buf], pf);
            bulk_g2s(dst + BUF_B_BYTES, idx + a0i, nbI, &bars[buf], pf);
        }
    };
    uint32_t phases = 0;
    auto wait_buf = [&](const Item& it, int buf) {
        if (!it.valid || it.b <= it.a) return;
        mbar_wait(&bars[buf], (phases >> buf) & 1u);
        phases ^= (1u << buf);
    };

    double bq[5][3], xq[5][3];
#pragma unroll
    for (int u = 0; u < 5; ++u) xq[u][0] = xq[u][1] = xq[u][2] = 0.0;
    double lamC[3] = {0.0, 0.0, 0.0}, lamN[3] = {0.0, 0.0, 0.0};
    auto load = [&](const Item& it, int buf, double (&lam)[3]) {
        if (!it.valid) return;
        item_load(wbuf + buf * BUF_BYTES, G, it, e, k, lane_on, bq, xq);
        if (MODE == 0 && it.last && lane < 9) {   // Lambda_T row for this segment's epilogue
            if (lamT != nullptr) {
                const double* L = lamT + 9 * (size_t)it.seg + 3 * (lane / 3);
                lam[0] = L[0]; lam[1] = L[1]; lam[2] = L[2];
            } else {   // raw gather (dual update input, bipgo.py:318): Lambda_T = I
                lam[0] = (lane < 3) ? 1.0 : 0.0; lam[1] = (lane >= 3 && lane < 6) ? 1.0 : 0.0; lam[2] = (lane >= 6) ? 1.0 : 0.0;
            }
        }
    };
    double acc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i] = 0.0;

    Item i0 = next_item(), i1 = next_item(), i2;
    issue(i0, 0);
    issue(i1, 1);
    wait_buf(i0, 0);
    load(i0, 0, lamC);
    int buf0 = 0;
    while (i0.valid) {
        item_fma(i0, bq, xq, acc);          // registers of item k are free after this
        i2 = next_item();
        issue(i2, buf0);                     // item k's stage is free (all its LDS fed the FMAs above)
        wait_buf(i1, buf0 ^ 1);
        load(i1, buf0 ^ 1, lamN);            // gathers of item k+1 in flight during the epilogue below
        if (i0.last) {                       // segment finished: combine the 30 private sums and emit
            int vidx;
            const double tot = reduce_scatter9(acc, lane, &vidx);
            const bool holder = ((lane & 1) == 0) && (vidx < 9);
            if (MODE == 0 && (VAR & 2)) {
                // total m sits in the lane pair inv[m] (see reduce_scatter9): output lane (a, j) fetches column j
                const int a = lane / 3, j = lane - 3 * a;
                const unsigned long long inv = 0xCA9854210ULL;   // m -> pair index
                const int jj = lane < 9 ? j : 0;
                const double z0 = shfl(tot, 2 * (int)((inv >> (4 * jj)) & 0xF));
                const double z1 = shfl(tot, 2 * (int)((inv >> (4 * (3 + jj))) & 0xF));
                const double z2 = shfl(tot, 2 * (int)((inv >> (4 * (6 + jj))) & 0xF));
                if (lane < 9) out[GSTRIDE * (size_t)i0.seg + 4 * a + j] = lamC[0] * z0 + lamC[1] * z1 + lamC[2] * z2;
            } else if (MODE == 0) {
                if (holder) scratch[vidx] = tot;
                __syncwarp();
                if (lane < 9) {
                    const int a = lane / 3, j = lane - 3 * a;
                    out[GSTRIDE * (size_t)i0.seg + 4 * a + j] =
                        lamC[0] * scratch[j] + lamC[1] * scratch[3 + j] + lamC[2] * scratch[6 + j];
                }
                __syncwarp();
            } else if (MODE == 2) {
                if (holder) atomicAdd(out + 9 * (size_t)__ldg(seg_node + i0.seg) + vidx, tot);
            } else {
                if (holder) out[9 * (size_t)i0.seg + vidx] = tot;
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] = 0.0;
        }
        if (MODE == 0 && i1.valid && i1.last) { lamC[0] = lamN[0]; lamC[1] = lamN[1]; lamC[2] = lamN[2]; }
        i0 = i1; i1 = i2;
        buf0 ^= 1;
    }
}

template <int MODE, int CTAS, int VAR>
__global__ void __launch_bounds__(PASS_THREADS, CTAS)
edge_pass_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node, const int* __restrict__ idx,
                 const double* __restrict__ B, const double* __restrict__ G, const double* __restrict__ lamT,
                 double* __restrict__ out, int n_seg, const double* __restrict__ skip_flag) {
    // speculatively enqueued launches (the host polls the eigen-solver's convergence flag one
    // iteration late) turn into no-ops once the flag is set
    if (skip_flag != nullptr && *skip_flag != 0.0) return;
    edge_pass_body<MODE, VAR>(seg_ptr, seg_node, idx, B, G, lamT, out, n_seg);
}

// Camera pass FUSED with the cross-rank sum of its result (edge-sharded multi-GPU runs): the
// per-tile atomics accumulate into this rank's peer window, and the kernel's epilogue is the
// one-shot all-reduce over NVLink peer memory of peer.cuh -- Y_c = sum over ranks, identical bits
// on every rank, no separate collective launch, no memset of Y.  Cooperative launch (the epilogue
// spins on the peers' flags, so all CTAs must be resident).
template <int CTAS>
__global__ void __launch_bounds__(PASS_THREADS, CTAS)
edge_pass_fused_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_node, const int* __restrict__ idx,
                       const double* __restrict__ B, const double* __restrict__ G, double* __restrict__ Y, long long n_y,
                       int n_seg, const double* __restrict__ skip_flag, PeerDev pd) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;   // identical on every rank (replicated, bitwise equal state)
    const unsigned long long epoch = pd.ctrl[pd.rank]->epoch + 1ull;
    double* mine = pd.buf[pd.rank] + (long long)(epoch & 1ull) * pd.cap;
    edge_pass_body<2, PASS_VAR_DEFAULT>(seg_ptr, seg_node, idx, B, G, nullptr, mine, n_seg);
    peer_publish(pd, epoch);
    peer_reduce(pd, epoch, Y, n_y);
}

// Y_c = sum of the camera's per-tile sums, windows ascending, tiles ascending inside a run (fixed order)
__global__ void tile_combine_kernel(const double* __restrict__ part, const int* __restrict__ tile_off, int64_t n_win, int64_t n_c,
                                    double* __restrict__ Y, const double* __restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * n_c) return;
    const int64_t c = i / 9;
    const int k = (int)(i - 9 * c);
    double acc = 0.0;
    for (int64_t w = 0; w < n_win; ++w) {
        const int t0 = __ldg(tile_off + w * n_c + c), t1 = __ldg(tile_off + w * n_c + c + 1);
        for (int t = t0; t < t1; ++t) acc += part[9 * (size_t)t + k];
    }
    Y[i] = acc;
}

// compact [n][9] -> padded [n][GSTRIDE] node blocks (the gather source layout)
__global__ void pad_blocks_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t n,
                                  const double* __restrict__ skip_flag) {
    if (skip_flag != nullptr && *skip_flag != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GSTRIDE * n) return;
    const int64_t node = i / GSTRIDE;
    const int c = (int)(i - GSTRIDE * node), row = c >> 2, col = c & 3;
    dst[i] = (row < 3 && col < 3) ? src[9 * node + 3 * row + col] : 0.0;
}

inline int launch_pad_blocks(const double* src, double* dst, int64_t n, cudaStream_t st, const double* skip_flag = nullptr) {
    if (n <= 0) return 0;
    pad_blocks_kernel<<<(int)((GSTRIDE * n + 255) / 256), 256, 0, st>>>(src, dst, n, skip_flag);
    VB_KERNEL_CHECK();
    return 0;
}

inline int pass_grid(int64_t n_segments, int ctas_per_sm) {
    const int64_t want = (n_segments + PASS_WARPS - 1) / PASS_WARPS;
    const int64_t cap = (int64_t)sm_count() * ctas_per_sm;   // persistent: exactly one resident wave
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int MODE, int VAR>
inline int launch_edge_pass_var(const int* seg_ptr, const int* seg_node, const int* idx, const double* B, const double* G,
                                const double* lamT, double* out, int64_t n_seg, cudaStream_t st, const double* skip_flag) {
    constexpr int CTAS = PASS_CTAS_PER_SM;   // 5 or 6 CTAs / SM (80-96 registers) measured slower, see profiles/
    static bool attr_set = false;
    if (!attr_set) {
        VB_CHECK(cudaFuncSetAttribute(edge_pass_kernel<MODE, CTAS, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, PASS_SMEM));
        attr_set = true;
    }
    edge_pass_kernel<MODE, CTAS, VAR><<<pass_grid(n_seg, CTAS), PASS_THREADS, PASS_SMEM, st>>>(seg_ptr, seg_node, idx, B, G, lamT, out, (int)n_seg, skip_flag);
    VB_KERNEL_CHECK();
    return 0;
}

template <int MODE>
inline int launch_edge_pass(const int* seg_ptr, const int* seg_node, const int* idx, const double* B, const double* G,
                            const double* lamT, double* out, int64_t n_seg, cudaStream_t st, const double* skip_flag) {
    // VICAN_B200_PASS_VAR: tuning variants of the kernel (profiles/r2_edge_pass.md); default = PASS_VAR_DEFAULT
    static const int var = getenv("VICAN_B200_PASS_VAR") ? atoi(getenv("VICAN_B200_PASS_VAR")) : PASS_VAR_DEFAULT;
    switch (var & 3) {
        case 0: return launch_edge_pass_var<MODE, 0>(seg_ptr, seg_node, idx, B, G, lamT, out, n_seg, st, skip_flag);
        case 1: return launch_edge_pass_var<MODE, 1>(seg_ptr, seg_node, idx, B, G, lamT, out, n_seg, st, skip_flag);
        case 2: return launch_edge_pass_var<MODE, 2>(seg_ptr, seg_node, idx, B, G, lamT, out, n_seg, st, skip_flag);
        default: return launch_edge_pass_var<MODE, 3>(seg_ptr, seg_node, idx, B, G, lamT, out, n_seg, st, skip_flag);
    }
}

// fused camera pass + cross-rank sum (see edge_pass_fused_kernel); Y needs no zeroing
inline int launch_pass_cam_fused(const int* tile_cam, const int* tile_start, const int* tidx, const double* B,
                                 const double* W12, double* Y, int64_t n_c, int64_t n_tiles, PeerCtx* peer, cudaStream_t st,
                                 const double* skip_flag = nullptr) {
    if (9 * n_c > peer->dev.cap) return 3;   // VB_STATUS_BAD_ARGUMENT
    constexpr int CTAS = PASS_CTAS_PER_SM;
    static bool attr_set = false;
    if (!attr_set) {
        VB_CHECK(cudaFuncSetAttribute(edge_pass_fused_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, PASS_SMEM));
        attr_set = true;
    }
    // every rank must take part even with no local tiles: at least one CTA
    const int grid = pass_grid(n_tiles < 1 ? 1 : n_tiles, CTAS);
    long long n_y = 9 * n_c;
    int n_seg = (int)n_tiles;
    void* args[] = {(void*)&tile_start, (void*)&tile_cam, (void*)&tidx, (void*)&B, (void*)&W12, (void*)&Y, (void*)&n_y,
                    (void*)&n_seg, (void*)&skip_flag, (void*)&peer->dev};
    VB_CHECK(cudaLaunchCooperativeKernel((void*)edge_pass_fused_kernel<CTAS>, dim3(grid), dim3(PASS_THREADS), args, PASS_SMEM, st));
    return 0;
}

// NOTE: idx must be readable up to index ((E+3)&~3)-1 and B up to edge ((E+1)&~1)-1 (the bulk
// copies are 16-byte granular); the ingestion allocates that padding.
// X12: padded gather source [n_c][GSTRIDE]; out12: padded [n_t][GSTRIDE].
inline int launch_pass_time(int mode, const int* rowptr, const int* cam, const double* B, const double* X12,
                            const double* lamT, double* out12, int64_t n_t, cudaStream_t st,
                            const double* skip_flag = nullptr) {
    if (n_t <= 0) return 0;
    // mode 1 (raw gather for the dual update) runs the same kernel with Lambda_T = I (lamT == nullptr): one
    // template instance to tune and profile; both variants measured the same 0.82-0.87 ms per 50 M-edge pass
    return launch_edge_pass<0>(rowptr, nullptr, cam, B, X12, mode == 0 ? lamT : nullptr, out12, n_t, st, skip_flag);
}

// tile_start carries a sentinel: tile_start[n_tiles] = E (tiles are contiguous).  W12 padded, Y compact.
// Two launches: per-tile sums into g->tile_part, then the per-camera combine (Y needs no zeroing).
inline int launch_pass_cam(const vb_graph* g, const double* W12, double* Y, cudaStream_t st, const double* skip_flag = nullptr) {
    if (g->n_tiles > 0) {
        int rc = launch_edge_pass<3>(g->tile_start, g->tile_cam, g->c_time, g->c_B, W12, nullptr, g->tile_part, g->n_tiles, st, skip_flag);
        if (rc) return rc;
    }
    tile_combine_kernel<<<(int)((9 * g->n_c + 255) / 256), 256, 0, st>>>(g->tile_part, g->tile_off, g->n_windows, g->n_c, Y, skip_flag);
    VB_KERNEL_CHECK();
    return 0;
}

}  // namespace vb
