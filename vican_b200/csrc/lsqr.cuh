// LSQR on the never-formed incidence system  J x = t~  (rows k_t (x_t - x_c) = k_t d_e), replaying
// scipy.sparse.linalg.lsqr as the reference calls it for lsqr_solver="direct"
// (vican/bipgo.py:479-480; scipy/sparse/linalg/_isolve/lsqr.py: Golub-Kahan bidiagonalisation,
// plane rotations via _sym_ortho, stopping tests test1/test2/test3).  Vector work runs in the
// kernels below; the scalar recurrences run in single-thread kernels so no scalar leaves the
// device except the stop flag.  All reductions are fixed-order (deterministic): LSQR amplifies
// 1e-16 perturbations of its inputs to ~1e-8 of the solution, so summation order is kept stable.
#pragma once
#include "translation.cuh"

namespace vb {

enum { LS_ALFA = 0, LS_BETA, LS_RHOBAR, LS_PHIBAR, LS_ANORM, LS_DDNORM, LS_XXNORM, LS_Z, LS_CS2, LS_SN2,
       LS_BNORM, LS_RHO, LS_T1, LS_T2, LS_THETA, LS_TAU, LS_PHI, LS_ISTOP, LS_ITN, LS_INV, LS_RES2, LS_XNORM,
       LS_NSCAL = 32 };

struct LsqrWork {
    double *u, *v_c, *v_t, *w_c, *w_t, *kt_sorted, *partial, *sc, *acc_c;
    int *row_cam, *row_time, *cam_rows, *cam_ptr;
    uint64_t *keys_a, *keys_b;
    int* vals_a;
    void* cub_tmp;
    size_t cub_bytes;
    int64_t bytes;
};

constexpr int LS_MAX_PARTIAL = 4096;

inline LsqrWork carve_lsqr(void* base, int64_t n_c, int64_t n_t, int64_t n_raw) {
    LsqrWork w;
    char* p = (char*)base;
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void* r = p + off;
        off += (bytes + 255) & ~(int64_t)255;
        return r;
    };
    w.u = (double*)take(8 * 3 * n_raw);
    w.v_c = (double*)take(8 * 3 * n_c); w.v_t = (double*)take(8 * 3 * n_t);
    w.w_c = (double*)take(8 * 3 * n_c); w.w_t = (double*)take(8 * 3 * n_t);
    w.kt_sorted = (double*)take(8 * n_raw);
    w.partial = (double*)take(8 * LS_MAX_PARTIAL);
    w.acc_c = (double*)take(8 * (3 * n_c + 8));
    w.sc = (double*)take(8 * LS_NSCAL);
    w.row_cam = (int*)take(4 * n_raw); w.row_time = (int*)take(4 * n_raw);
    w.cam_rows = (int*)take(4 * n_raw); w.cam_ptr = (int*)take(4 * (n_c + 1));
    w.keys_a = (uint64_t*)take(8 * n_raw); w.keys_b = (uint64_t*)take(8 * n_raw);
    w.vals_a = (int*)take(4 * n_raw);
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int*)nullptr,
                                    (int*)nullptr, (int)n_raw, 0, 64);
    w.cub_bytes = a + 1024;
    w.cub_tmp = take((int64_t)w.cub_bytes);
    w.bytes = off;
    return w;
}

__global__ void lsqr_rows_kernel(const int* __restrict__ raw_perm, const int* __restrict__ raw_pair, const int* __restrict__ t_cam,
                                 const int* __restrict__ t_time, const double* __restrict__ k_t, const double* __restrict__ d_sorted,
                                 int64_t n_raw, int* __restrict__ row_cam, int* __restrict__ row_time, double* __restrict__ kt_sorted,
                                 double* __restrict__ u, uint64_t* __restrict__ keys, int* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_raw) return;
    const int p = raw_pair[i];
    const int c = t_cam[p];
    row_cam[i] = c; row_time[i] = t_time[p];
    const double kt = k_t[raw_perm[i]];
    kt_sorted[i] = kt;
#pragma unroll
    for (int k = 0; k < 3; ++k) u[3 * i + k] = kt * d_sorted[3 * i + k];   // u = b = t~  (bipgo.py:454-461)
    keys[i] = (uint64_t)c;
    vals[i] = (int)i;
}

// deterministic two-level sum of squares of a vector (fixed block order)
__global__ void sumsq_partial_kernel(const double* __restrict__ a, int64_t n, double scale, double* __restrict__ partial) {
    __shared__ double sm[TR_THREADS / 32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = a[i] * scale;
        s += v * v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < TR_THREADS / 32; ++w) t += sm[w];
        partial[blockIdx.x] = t;
    }
}
inline int sumsq_grid(int64_t n) { int g = tr_grid(n); return g > 1024 ? 1024 : (g < 1 ? 1 : g); }

__device__ __forceinline__ double sum_partials(const double* partial, int a, int b) {
    double s = 0.0;
    for (int i = a; i < b; ++i) s += partial[i];
    return s;
}

// u = k_t (v_t - v_c) - alfa u        (A.matvec(v) - alfa * u)
__global__ void lsqr_u_kernel(const int* __restrict__ row_cam, const int* __restrict__ row_time, const double* __restrict__ kt,
                              const double* __restrict__ v_c, const double* __restrict__ v_t, double* __restrict__ u, int64_t n_raw,
                              const double* sc) {
    if (sc[LS_ISTOP] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_raw) return;
    const double alfa = sc[LS_ALFA];
    const int64_t c = row_cam[i], t = row_time[i];
    const double k = kt[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) u[3 * i + j] = k * (v_t[3 * t + j] - v_c[3 * c + j]) - alfa * u[3 * i + j];
}

// time side of v = A^T u - beta v ; normalises u in place (u <- u / beta) on the way.
// rows of one time node are contiguous in the sorted raw list: [row_start[t], row_start[t+1])
__global__ void lsqr_vt_kernel(const int* __restrict__ t_rowptr, const int* __restrict__ pair_start, const double* __restrict__ kt,
                               double* __restrict__ u, double* __restrict__ v_t, int64_t n_t, const double* sc, int init) {
    if (sc[LS_ISTOP] != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= n_t) return;
    const double inv = sc[LS_INV], beta = sc[LS_BETA];
    const int s = pair_start[t_rowptr[t]], e = pair_start[t_rowptr[t + 1]];
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = s + lane; i < e; i += 32) {
        const double k = kt[i];
        const double u0 = u[3 * i] * inv, u1 = u[3 * i + 1] * inv, u2 = u[3 * i + 2] * inv;
        u[3 * i] = u0; u[3 * i + 1] = u1; u[3 * i + 2] = u2;
        a0 += k * u0; a1 += k * u1; a2 += k * u2;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        if (init) { v_t[3 * t] = a0; v_t[3 * t + 1] = a1; v_t[3 * t + 2] = a2; }
        else { v_t[3 * t] = a0 - beta * v_t[3 * t]; v_t[3 * t + 1] = a1 - beta * v_t[3 * t + 1]; v_t[3 * t + 2] = a2 - beta * v_t[3 * t + 2]; }
    }
}

// camera side (u already normalised): warp per camera over its rows in ascending row order.
// acc_out != nullptr (edge-sharded runs): only the local row sums are written there; they are summed over the
// ranks and lsqr_vc_finish_kernel applies  v_c = sum - beta v_c.
__global__ void lsqr_vc_kernel(const int* __restrict__ cam_ptr, const int* __restrict__ cam_rows, const double* __restrict__ kt,
                               const double* __restrict__ u, double* __restrict__ v_c, int64_t n_c, const double* sc, int init,
                               double* __restrict__ acc_out) {
    if (sc[LS_ISTOP] != 0.0) return;
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_c) return;
    const double beta = sc[LS_BETA];
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = cam_ptr[c] + lane; i < cam_ptr[c + 1]; i += 32) {
        const int64_t r = cam_rows[i];
        const double k = kt[r];
        a0 -= k * u[3 * r]; a1 -= k * u[3 * r + 1]; a2 -= k * u[3 * r + 2];
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        if (acc_out) { acc_out[3 * c] = a0; acc_out[3 * c + 1] = a1; acc_out[3 * c + 2] = a2; }
        else if (init) { v_c[3 * c] = a0; v_c[3 * c + 1] = a1; v_c[3 * c + 2] = a2; }
        else { v_c[3 * c] = a0 - beta * v_c[3 * c]; v_c[3 * c + 1] = a1 - beta * v_c[3 * c + 1]; v_c[3 * c + 2] = a2 - beta * v_c[3 * c + 2]; }
    }
}

__global__ void lsqr_vc_finish_kernel(const double* __restrict__ acc, double* __restrict__ v_c, int64_t n3, const double* sc, int init) {
    if (sc[LS_ISTOP] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    v_c[i] = init ? acc[i] : acc[i] - sc[LS_BETA] * v_c[i];
}

// edge-sharded runs: fixed-order sum of a partial range into one slot / copy of one slot (single thread)
__global__ void lsqr_collapse_kernel(const double* partial, int nb, double* dst) { *dst = sum_partials(partial, 0, nb); }
__global__ void lsqr_set_kernel(double* dst, const double* src) { *dst = *src; }

// scalar step 1: beta = ||u||, anorm update.  init: beta = bnorm.
__global__ void lsqr_s1_kernel(double* sc, const double* partial, int nb, int init) {
    if (sc[LS_ISTOP] != 0.0) return;
    const double beta = sqrt(sum_partials(partial, 0, nb));
    sc[LS_BETA] = beta;
    sc[LS_INV] = beta > 0.0 ? 1.0 / beta : 0.0;
    if (init) { sc[LS_BNORM] = beta; }
    else {
        sc[LS_ITN] += 1.0;
        // anorm is updated once alfa of THIS step is the old one: anorm = sqrt(anorm^2 + alfa^2 + beta^2)
        const double an = sc[LS_ANORM], al = sc[LS_ALFA];
        sc[LS_ANORM] = sqrt(an * an + al * al + beta * beta);
    }
}

__device__ __forceinline__ void sym_ortho(double a, double b, double& c, double& s, double& r) {
    auto sgn = [](double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); };
    if (b == 0.0) { c = sgn(a); s = 0.0; r = fabs(a); }
    else if (a == 0.0) { c = 0.0; s = sgn(b); r = fabs(b); }
    else if (fabs(b) > fabs(a)) { const double tau = a / b; s = sgn(b) / sqrt(1.0 + tau * tau); c = s * tau; r = b / s; }
    else { const double tau = b / a; c = sgn(a) / sqrt(1.0 + tau * tau); s = c * tau; r = a / c; }
}

// scalar step 2: alfa = ||v||; plane rotation; coefficients for the x/w update.
__global__ void lsqr_s2_kernel(double* sc, const double* partial, int nb_c, int nb_t, int init) {
    if (sc[LS_ISTOP] != 0.0) return;
    const double alfa = sqrt(sum_partials(partial, 0, nb_c) + sum_partials(partial, 1024, 1024 + nb_t));
    sc[LS_ALFA] = alfa;
    sc[LS_INV] = alfa > 0.0 ? 1.0 / alfa : 0.0;
    if (init) {
        sc[LS_RHOBAR] = alfa; sc[LS_PHIBAR] = sc[LS_BETA];
        sc[LS_ANORM] = 0.0; sc[LS_DDNORM] = 0.0; sc[LS_XXNORM] = 0.0; sc[LS_Z] = 0.0; sc[LS_CS2] = -1.0; sc[LS_SN2] = 0.0;
        sc[LS_RES2] = 0.0; sc[LS_ITN] = 0.0;
        if (alfa * sc[LS_BETA] == 0.0) sc[LS_ISTOP] = -1.0;   // "exact solution is x = 0"
        return;
    }
    double cs, sn, rho;
    sym_ortho(sc[LS_RHOBAR], sc[LS_BETA], cs, sn, rho);
    const double theta = sn * alfa;
    sc[LS_RHOBAR] = -cs * alfa;
    const double phi = cs * sc[LS_PHIBAR];
    sc[LS_PHIBAR] = sn * sc[LS_PHIBAR];
    sc[LS_TAU] = sn * phi;
    sc[LS_RHO] = rho; sc[LS_THETA] = theta; sc[LS_PHI] = phi;
    sc[LS_T1] = phi / rho;
    sc[LS_T2] = -theta / rho;
}

// v <- v / alfa ; x += t1 w ; w = v + t2 w   (init: w = v, x = 0)
__global__ void lsqr_x_kernel(double* __restrict__ v, double* __restrict__ w, double* __restrict__ x, int64_t n3, const double* sc, int init) {
    if (sc[LS_ISTOP] != 0.0) return;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const double vv = v[i] * sc[LS_INV];
    v[i] = vv;
    if (init) { w[i] = vv; x[i] = 0.0; }
    else {
        const double wi = w[i];
        x[i] += sc[LS_T1] * wi;
        w[i] = vv + sc[LS_T2] * wi;
    }
}

// scalar step 3: ddnorm, xnorm estimate, stopping tests (lsqr.py:470-545)
__global__ void lsqr_s3_kernel(double* sc, const double* partial, int nb_c, int nb_t, double atol, double btol, double conlim,
                               double iter_lim) {
    if (sc[LS_ISTOP] != 0.0) return;
    const double eps = 2.220446049250313e-16;
    const double rho = sc[LS_RHO], theta = sc[LS_THETA], phi = sc[LS_PHI];
    // ||dk||^2 = ||w_old||^2 / rho^2 ; partials hold ||w_old||^2 (taken before the update)
    const double wn2 = sum_partials(partial, 0, nb_c) + sum_partials(partial, 1024, 1024 + nb_t);
    const double dkn = sqrt(wn2) * fabs(1.0 / rho);
    sc[LS_DDNORM] += dkn * dkn;
    const double delta = sc[LS_SN2] * rho;
    const double gambar = -sc[LS_CS2] * rho;
    const double rhs = phi - delta * sc[LS_Z];
    const double zbar = rhs / gambar;
    const double xnorm = sqrt(sc[LS_XXNORM] + zbar * zbar);
    const double gamma = sqrt(gambar * gambar + theta * theta);
    sc[LS_CS2] = gambar / gamma;
    sc[LS_SN2] = theta / gamma;
    const double z = rhs / gamma;
    sc[LS_Z] = z;
    sc[LS_XXNORM] += z * z;
    sc[LS_XNORM] = xnorm;
    const double anorm = sc[LS_ANORM], bnorm = sc[LS_BNORM];
    const double acond = anorm * sqrt(sc[LS_DDNORM]);
    const double res1 = sc[LS_PHIBAR] * sc[LS_PHIBAR];
    const double rnorm = sqrt(res1 + sc[LS_RES2]);
    const double arnorm = sc[LS_ALFA] * fabs(sc[LS_TAU]);
    const double test1 = rnorm / bnorm;
    const double test2 = arnorm / (anorm * rnorm + eps);
    const double test3 = 1.0 / (acond + eps);
    const double t1 = test1 / (1.0 + anorm * xnorm / bnorm);
    const double rtol = btol + atol * anorm * xnorm / bnorm;
    const double ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;
    double istop = 0.0;
    if (sc[LS_ITN] >= iter_lim) istop = 7.0;
    if (1.0 + test3 <= 1.0) istop = 6.0;
    if (1.0 + test2 <= 1.0) istop = 5.0;
    if (1.0 + t1 <= 1.0) istop = 4.0;
    if (test3 <= ctol) istop = 3.0;
    if (test2 <= atol) istop = 2.0;
    if (test1 <= rtol) istop = 1.0;
    sc[LS_ISTOP] = istop;
}

__global__ void lsqr_clear_kernel(double* sc) {
    for (int i = 0; i < LS_NSCAL; ++i) sc[i] = 0.0;
}

}  // namespace vb
