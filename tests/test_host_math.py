"""CPU unit tests of the header-only device math (compiled with the host compiler):
3x3 Jacobi SVD factors vs np.linalg.svd, 9x9 Rayleigh-Ritz vs numpy."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "host", "_host_math.so")
SRC = os.path.join(HERE, "host", "host_math_shim.cpp")
HDRS = [os.path.join(HERE, "..", "vican_b200", "csrc", h) for h in ("mat3.cuh", "dense_small.cuh")]


@pytest.fixture(scope="module")
def lib():
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO])
    return ctypes.CDLL(SO)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_factors(M):
    U, S, Vt = np.linalg.svd(M)
    d = np.linalg.det(U @ Vt)
    D = np.zeros_like(M)
    D[:, 0, 0] = D[:, 1, 1] = 1
    D[:, 2, 2] = d
    Ut = np.transpose(U, (0, 2, 1))
    return U @ D @ Vt, (U * S[:, None, :]) @ Ut, (U * (1 / S)[:, None, :]) @ Ut, S


def test_svd3_factors_random(lib):
    rng = np.random.default_rng(0)
    n = 20000
    M = rng.standard_normal((n, 3, 3))
    M[: n // 4] *= rng.uniform(1e-3, 1e3, (n // 4, 1, 1))
    # sums of noisy rotations (what the solver feeds it)
    from vican_b200.synthetic import random_rotations, so3_exp
    R = random_rotations(rng, n // 4)
    M[n // 4: n // 2] = sum(R @ so3_exp(rng.normal(0, 0.05, (n // 4, 3))) * rng.uniform(0.5, 1.5) for _ in range(7))
    M[n // 2: n // 2 + 100, :, 0] *= -1  # negative determinants
    rot = np.zeros_like(M); sp = np.zeros_like(M); si = np.zeros_like(M)
    lib.h_svd3_factors(P(M), P(rot), P(sp), P(si), ctypes.c_int64(n))
    r0, p0, i0, S = ref_factors(M)
    cond = S[:, 0] / S[:, 2]
    scale = S[:, 0]
    assert np.all(np.abs(rot - r0).max(axis=(1, 2)) < 1e-13 * cond)
    assert np.all(np.abs(sp - p0).max(axis=(1, 2)) < 1e-13 * scale * cond)
    assert np.all(np.abs(si - i0).max(axis=(1, 2)) < 1e-13 * cond * cond / S[:, 2])
    assert np.allclose(np.linalg.det(rot), 1.0, atol=1e-12)


def test_node_factors_newton_polar_matches_svd(lib):
    """node_factors = scaled-Newton polar route with SVD fallback: same three factors as LAPACK."""
    rng = np.random.default_rng(3)
    n = 30000
    from vican_b200.synthetic import random_rotations, so3_exp
    M = rng.standard_normal((n, 3, 3)) * rng.uniform(1e-2, 1e2, (n, 1, 1))
    R = random_rotations(rng, n // 2)
    M[: n // 2] = sum(R @ so3_exp(rng.normal(0, 0.05, (n // 2, 3))) * rng.uniform(0.5, 1.5, (n // 2, 1, 1)) for _ in range(20))
    M[n // 2: n // 2 + 200, :, 0] *= -1                      # det < 0 -> SVD fallback (det fix)
    M[n // 2 + 200: n // 2 + 300, :, 2] = M[n // 2 + 200: n // 2 + 300, :, 0] * (1 + 1e-11)   # near singular
    rot = np.zeros_like(M); sp = np.zeros_like(M); si = np.zeros_like(M)
    lib.h_node_factors(P(M), P(rot), P(sp), P(si), ctypes.c_int64(n))
    r0, p0, i0, S = ref_factors(M)
    cond = S[:, 0] / S[:, 2]
    good = cond < 1e8
    assert np.all(np.abs(rot - r0).max(axis=(1, 2))[good] < 1e-13 * cond[good])
    assert np.all(np.abs(sp - p0).max(axis=(1, 2))[good] < 1e-13 * (S[:, 0] * cond)[good])
    assert np.all(np.abs(si - i0).max(axis=(1, 2))[good] < 1e-13 * (cond * cond / S[:, 2])[good])
    assert np.allclose(np.linalg.det(rot[good]), 1.0, atol=1e-12)
    # consistent rotations (what the solver feeds): essentially exact
    assert np.abs(rot[: n // 2] - r0[: n // 2]).max() < 1e-13


def test_svd3_singular_values_and_orthogonality(lib):
    rng = np.random.default_rng(1)
    n = 5000
    M = rng.standard_normal((n, 3, 3))
    M[:50, :, 2] = M[:50, :, 0]          # rank 2
    M[50:60] = 0.0                        # rank 0
    M[60:70, :, 1:] = 0.0                 # rank 1
    U = np.zeros_like(M); V = np.zeros_like(M); S = np.zeros((n, 3))
    lib.h_svd3(P(M), P(U), P(S), P(V), ctypes.c_int64(n))
    S0 = np.linalg.svd(M, compute_uv=False)
    assert np.abs(S - S0).max() < 1e-13 * max(1.0, S0.max())
    I = np.eye(3)[None]
    assert np.abs(np.transpose(U, (0, 2, 1)) @ U - I).max() < 1e-13
    assert np.abs(np.transpose(V, (0, 2, 1)) @ V - I).max() < 1e-13
    rec = (U * S[:, None, :]) @ np.transpose(V, (0, 2, 1))
    assert np.abs(rec - M).max() < 1e-13 * max(1.0, S0.max())


def test_inv3(lib):
    rng = np.random.default_rng(2)
    A = rng.standard_normal((1000, 3, 3))
    I = np.zeros_like(A)
    lib.h_inv3(P(A), P(I), ctypes.c_int64(1000))
    err = np.abs(A @ I - np.eye(3)).max(axis=(1, 2))
    assert np.all(err < 1e-12 * np.linalg.cond(A))


@pytest.mark.parametrize("n", [3, 6, 9])
def test_jacobi_eig(lib, n):
    rng = np.random.default_rng(n)
    for trial in range(50):
        A = rng.standard_normal((n, n))
        A = A + A.T
        if trial % 3 == 0:
            A[0, 0] += 1e6
        if trial % 5 == 0:
            A[:, 1] = A[1, :] = 0
            A[1, 1] = 1e30
        A0 = A.copy()
        Q = np.zeros((n, n)); lam = np.zeros(n)
        lib.h_jacobi(n, P(A), P(Q), P(lam))
        w = np.linalg.eigvalsh(A0)
        sc = np.abs(w).max()
        assert np.abs(lam - w).max() < 1e-13 * sc
        assert np.abs(Q.T @ Q - np.eye(n)).max() < 1e-13
        assert np.abs(A0 @ Q - Q * lam[None]).max() < 1e-12 * sc


def test_svqb3(lib):
    rng = np.random.default_rng(5)
    for trial in range(30):
        W = rng.standard_normal((50, 3)) * rng.uniform(1e-6, 1e3, (1, 3))
        if trial % 4 == 0:
            W[:, 2] = W[:, 0] * 2 + 1e-9 * W[:, 2]   # nearly dependent -> dropped
        G = W.T @ W
        T = np.zeros((3, 3)); act = np.zeros(3, np.int32)
        lib.h_svqb3(P(G), P(T), P(act), ctypes.c_double(1e-12))
        Wn = W @ T
        k = act.astype(bool)
        if trial % 4 == 0:
            assert k.sum() == 2
        assert np.abs(Wn[:, k].T @ Wn[:, k] - np.eye(k.sum())).max() < 1e-9
        assert (not (~k).any()) or np.abs(Wn[:, ~k]).max() == 0


@pytest.mark.parametrize("fn", ["h_ritz9", "h_ritz9_coop"])
def test_ritz9_matches_numpy(lib, fn):
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = 40
        A = rng.standard_normal((n, n)); A = A + A.T
        S = rng.standard_normal((n, 9))
        S, _ = np.linalg.qr(S)
        S = S @ (np.eye(9) + 1e-3 * rng.standard_normal((9, 9)))   # slightly non-orthonormal basis
        act = np.ones(9, np.int32)
        if trial % 3 == 1:
            act[6:] = 0
        if trial % 3 == 2:
            act[3:] = 0
        S[:, act == 0] = 0
        G = S.T @ A @ S
        M = S.T @ S
        C = np.zeros((9, 3)); Cp = np.zeros((9, 3)); th = np.zeros(3); actP = np.zeros(3, np.int32)
        getattr(lib, fn)(P(G), P(M), P(act), P(C), P(Cp), P(th), P(actP))
        k = act.astype(bool)
        import scipy.linalg as sl
        w, v = sl.eigh(G[np.ix_(k, k)], M[np.ix_(k, k)])
        assert np.abs(th - w[:3]).max() < 1e-11 * np.abs(w).max()
        X = S @ C
        assert np.abs(X.T @ X - np.eye(3)).max() < 1e-11
        assert np.abs(X.T @ A @ X - np.diag(th)).max() < 1e-10 * np.abs(w).max()
        Pn = S @ Cp
        ka = actP.astype(bool)
        if k.sum() > 3:
            assert ka.sum() == 3
            assert np.abs(Pn.T @ Pn - np.eye(3)).max() < 1e-10
            assert np.abs(X.T @ Pn).max() < 1e-10
            # span[X', P'] contains the X-part-removed update directions
            Z = C.copy(); Z[:3] = 0
            Zs = S @ Z
            Bs = np.concatenate([X, Pn], 1)
            res = Zs - Bs @ (Bs.T @ Zs)
            assert np.abs(res).max() < 1e-9
        else:
            assert ka.sum() == 0
