// Host build of the header-only device math (mat3.cuh, dense_small.cuh) so that it can
// be unit-tested against numpy on a machine without a GPU.  Test infrastructure only.
#include "../../vican_b200/csrc/dense_small.cuh"
#include <cstdint>

extern "C" {
void h_svd3(const double* M, double* U, double* S, double* V, int64_t n) {
    for (int64_t i = 0; i < n; ++i) vb::svd3(M + 9 * i, U + 9 * i, S + 3 * i, V + 9 * i);
}
void h_svd3_factors(const double* M, double* rot, double* spos, double* sinv, int64_t n) {
    for (int64_t i = 0; i < n; ++i) vb::svd3_factors(M + 9 * i, rot + 9 * i, spos + 9 * i, sinv + 9 * i);
}
void h_node_factors(const double* M, double* rot, double* spos, double* sinv, int64_t n) {
    for (int64_t i = 0; i < n; ++i) vb::node_factors(M + 9 * i, rot + 9 * i, spos + 9 * i, sinv + 9 * i);
}
void h_inv3(const double* A, double* I, int64_t n) {
    for (int64_t i = 0; i < n; ++i) vb::inv3(A + 9 * i, I + 9 * i);
}
void h_jacobi(int n, double* A, double* Q, double* lam) { vb::jacobi_eig_sym(n, A, Q, lam); }
void h_svqb3(const double* G, double* T, int* act, double tol) { vb::svqb3(G, T, act, tol); }
void h_ritz9(const double* G, const double* M, const int* act, double* C, double* Cp, double* theta, int* actP) {
    double work[4 * 81];
    vb::ritz9(G, M, act, C, Cp, theta, actP, work);
}
void h_ritz9_coop(const double* G, const double* M, const int* act, double* C, double* Cp, double* theta, int* actP) {
    double work[5 * 81 + 64];
    int iwork[48];
    vb::ritz9_coop(G, M, act, C, Cp, theta, actP, work, iwork, 0, 1);
}
}
