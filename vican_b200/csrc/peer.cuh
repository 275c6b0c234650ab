// One-shot all-reduce over NVLink peer memory (edge-sharded multi-GPU runs).
//
// Every rank owns one cudaMalloc'ed "window" that all other ranks of the box map through CUDA
// IPC:  two exchange buffers (parity of the call number) + one flag per peer + counters.
// An all-reduce of the camera-side accumulator (n_c x 9 doubles = 720 KB at cfg4) is then
//   1. my partial sums land in MY buffer[parity]           (written by my own kernel)
//   2. the last CTA to finish publishes  flag[me] = epoch  into every peer's window
//      (fence.sys + st.release.sys over NVLink)
//   3. every CTA waits until all flags in its OWN window reached the epoch (relaxed polls, one fence),
//      then sums its slice of the 8 partials straight out of the peers' windows in rank order
//      -> the result is bitwise identical on all ranks (the replicated LOBPCG / SVD steps
//      depend on that) and no data is ever staged or copied twice.
// The other-parity buffer of my window is cleared in step 3 for the next call, which is safe
// because every peer finished reading it before it published the flag I just waited for.
//
// Two users: peer_allreduce_kernel (generic vb_allreduce_fn for small vectors: CG, degrees) and
// edge_pass_fused_kernel in passes.cuh, where steps 2-3 are the EPILOGUE of the camera pass itself
// (the per-tile atomics accumulate directly into the window).  Both are launched cooperatively:
// the spin in step 3 requires all CTAs of the grid to be resident.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace vb {

constexpr int PEER_MAX = 8;
constexpr int PEER_AR_THREADS = 512;
constexpr int PEER_AR_CTAS = 128;

struct PeerCtrl {            // lives at the start of every window
    unsigned long long flags[PEER_MAX];   // flags[r]: latest epoch rank r published (written by rank r)
    unsigned long long epoch;             // calls completed by this rank
    unsigned int done;                    // CTAs that finished step 1
    unsigned int done2;                   // CTAs that finished step 3
    unsigned int timeouts;                // waits that gave up (a peer never arrived): results are invalid
    unsigned int pad;
    unsigned long long dirty[2];          // doubles of buffer 0 / 1 that are not zero (set by the call that used it)
    unsigned long long stamp[4];          // globaltimer (ns) of the last call: publish, flags seen (CTA 0), sums done (CTA 0), end
};
__device__ __forceinline__ unsigned long long peer_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr long long PEER_SPIN_LIMIT = 20000000000ll;   // clock64 ticks (~10 s): a lost peer becomes an error, not a hung GPU
constexpr size_t PEER_CTRL_BYTES = 256;   // >= sizeof(PeerCtrl), keeps the buffers 256-byte aligned

struct PeerDev {             // kernel argument
    int rank, world;
    long long cap;                        // doubles per buffer
    PeerCtrl* ctrl[PEER_MAX];             // every rank's window (ctrl[rank] = mine)
    double* buf[PEER_MAX];                // buffer 0 of every window; buffer 1 = buf + cap
};

struct PeerCtx {
    PeerDev dev;
    void* base = nullptr;
    void* opened[PEER_MAX] = {nullptr};
    size_t bytes = 0;
    const double* skip = nullptr;   // see PeerSkipScope
};

// While alive, the generic all-reduces of this context return at their first instruction when *flag != 0.  The flag
// must hold the same value on every rank at every call (it does for the CG's done flag: it is computed from sums
// that are bitwise identical on all ranks), so either all ranks exchange or none does and the epochs stay in step.
// Iterations a host loop enqueued past convergence then cost a launch, not an NVLink round trip.
struct PeerSkipScope {
    PeerCtx* ctx;
    PeerSkipScope(PeerCtx* c, const double* flag) : ctx(c) { if (ctx) ctx->skip = flag; }
    ~PeerSkipScope() { if (ctx) ctx->skip = nullptr; }
    PeerSkipScope(const PeerSkipScope&) = delete;
    PeerSkipScope& operator=(const PeerSkipScope&) = delete;
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// polling load: RELAXED (an ld.acquire.sys compiles to LDG + CCTL.IVALL, i.e. every poll iteration
// would flush the SM's L1 under the CTAs that are still gathering); the acquire is one fence afterwards
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// __threadfence() / __threadfence_system() are fence.sc (MEMBAR.SC: totally ordered, expensive when
// hundreds of CTAs issue them together); release / acquire patterns only need fence.acq_rel
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ double2 ld_peer2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}

// Step 2 (called by ALL threads of a CTA after its last write to the window): returns when this
// CTA may proceed to the wait.  The last CTA of the grid publishes the epoch to every peer.
__device__ __forceinline__ void peer_publish(const PeerDev& pd, unsigned long long epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_acq_rel_gpu();   // cumulative: covers the whole CTA's writes (ordered before this thread by the barrier)
        PeerCtrl* me = pd.ctrl[pd.rank];
        const unsigned int old = atomicAdd(&me->done, 1u);
        if (old == gridDim.x - 1) {
            // ONE system-scope fence orders every CTA's window writes (made visible to this thread through
            // the counter) before the flag stores; the stores themselves are relaxed and pipeline over NVLink
            // (a st.release per peer would serialise 8 fence + round-trip pairs)
            me->stamp[0] = peer_now();
            fence_acq_rel_sys();
#pragma unroll
            for (int r = 0; r < PEER_MAX; ++r)
                if (r < pd.world) st_relaxed_sys(&pd.ctrl[r]->flags[pd.rank], epoch);
        }
    }
}

// Step 3: wait for all ranks, then out[i] = sum_r window_r[parity][i] for this CTA's share of
// [0, count), clear my other-parity buffer, and let the last CTA advance the epoch.
__device__ __forceinline__ void peer_reduce(const PeerDev& pd, unsigned long long epoch, double* __restrict__ out, long long count) {
    PeerCtrl* me = pd.ctrl[pd.rank];
    const int opar = (int)((epoch + 1ull) & 1ull);
    const long long dirty_other = (long long)me->dirty[opar];   // written by the previous call's last CTA
    if (threadIdx.x == 0) {                     // one polling thread per CTA, with back-off
        const long long t0 = clock64();
        for (int r = 0; r < pd.world; ++r) {
            while (ld_relaxed_sys(&me->flags[r]) < epoch) {
                __nanosleep(200);
                if (clock64() - t0 > PEER_SPIN_LIMIT) { atomicAdd(&me->timeouts, 1u); break; }
            }
        }
        fence_acq_rel_sys();                    // acquire: the peers' window writes are ordered before their flags
        if (blockIdx.x == 0) me->stamp[1] = peer_now();
    }
    __syncthreads();
    const long long par = (long long)(epoch & 1ull) * pd.cap;
    const long long npair = (count + 1) >> 1;                 // buffers are padded to an even length
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += stride) {
        // all peer loads in flight together (one NVLink round trip, not `world` of them); summed in rank order
        double2 v[PEER_MAX];
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < pd.world) v[r] = ld_peer2(pd.buf[r] + par + 2 * i);
        double2 s = make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < PEER_MAX; ++r)
            if (r < pd.world) { s.x += v[r].x; s.y += v[r].y; }
        out[2 * i] = s.x;
        if (2 * i + 1 < count) out[2 * i + 1] = s.y;
    }
    // the next call may be a fused camera pass that ACCUMULATES into the other buffer: clear what the
    // call before this one left there (nobody reads it any more, see the header)
    double* other = pd.buf[pd.rank] + (long long)opar * pd.cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < dirty_other; i += stride) other[i] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) me->stamp[2] = peer_now();
        fence_acq_rel_gpu();
        const unsigned int old = atomicAdd(&me->done2, 1u);
        if (old == gridDim.x - 1) {
            me->done = 0u; me->done2 = 0u;
            me->dirty[opar] = 0ull;
            me->dirty[opar ^ 1] = (unsigned long long)count;
            me->stamp[3] = peer_now();
            fence_acq_rel_gpu();
            me->epoch = epoch;
        }
    }
}

// generic in-place all-reduce of a small vector (count <= cap)
__global__ void __launch_bounds__(PEER_AR_THREADS) peer_allreduce_kernel(PeerDev pd, double* __restrict__ data, long long count,
                                                                             const double* __restrict__ skip) {
    if (skip != nullptr && *skip != 0.0) return;
    const unsigned long long epoch = pd.ctrl[pd.rank]->epoch + 1ull;   // advanced by the previous call's last CTA
    double* mine = pd.buf[pd.rank] + (long long)(epoch & 1ull) * pd.cap;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) mine[i] = data[i];
    peer_publish(pd, epoch);
    peer_reduce(pd, epoch, data, count);
}

inline int launch_peer_allreduce(PeerCtx* ctx, double* data, int64_t count, cudaStream_t st) {
    if (count <= 0) return 0;
    if (count > ctx->dev.cap) return 3;   // VB_STATUS_BAD_ARGUMENT
    int grid = (int)((count + PEER_AR_THREADS - 1) / PEER_AR_THREADS);
    if (grid < 16) grid = 16;             // even a 3-scalar call may have to clear a camera accumulator (see peer_reduce)
    if (grid > PEER_AR_CTAS) grid = PEER_AR_CTAS;
    long long cnt = count;
    const double* skip = ctx->skip;
    void* args[] = {(void*)&ctx->dev, (void*)&data, (void*)&cnt, (void*)&skip};
    VB_CHECK(cudaLaunchCooperativeKernel((void*)peer_allreduce_kernel, dim3(grid), dim3(PEER_AR_THREADS), args, 0, st));
    return 0;
}

}  // namespace vb
