"""numpy MODEL of the device algorithm (test infrastructure, NOT product code).

The CUDA path does not form the power-graph Laplacian and does not call ARPACK:
it applies L = Lambda_C - P Lambda_T P^T matrix-free in two edge passes (time-sorted
gather, camera-sorted gather) and finds the 3-dimensional invariant subspace with a
width-3 LOBPCG (block-Jacobi preconditioner Lambda_C^-1, warm start, 9x9
Rayleigh-Ritz).  This file states that algorithm in numpy, step for step as the
kernels in ``vican_b200/csrc`` execute it, so that its convergence and its parity
with the reference-faithful oracle (``vican_oracle.so3sync``) can be checked on CPU
and so that GPU intermediates can be compared against it.
"""
from __future__ import annotations

import numpy as np

BIG = 1e30


def pass_time(pc, pt, B, X, n_t):
    """Z_t = sum_{e in t} B_e^T X_{c_e}   (kernel vb_pass_time)."""
    Z = np.zeros((n_t, 3, 3))
    np.add.at(Z, pt, np.transpose(B, (0, 2, 1)) @ X[pc])
    return Z


def pass_cam(pc, pt, B, W, n_c):
    """Y_c = sum_{e in c} B_e W_{t_e}     (kernel vb_pass_cam)."""
    Y = np.zeros((n_c, 3, 3))
    np.add.at(Y, pc, B @ W[pt])
    return Y


def sym6_to_mat(s):
    return s


def svd_factors(M):
    U, S, Vt = np.linalg.svd(M)
    d = np.linalg.det(U @ Vt)
    D = np.zeros_like(M)
    D[:, 0, 0] = D[:, 1, 1] = 1.0
    D[:, 2, 2] = d
    rot = U @ D @ Vt
    Ut = np.transpose(U, (0, 2, 1))
    return rot, (U * S[:, None, :]) @ Ut, (U * (1.0 / S)[:, None, :]) @ Ut


class LobpcgStats:
    def __init__(self):
        self.applies = []
        self.resid = []
        self.theta = []


def _svqb(W, drop_tol=1e-12):
    """Orthonormalise the columns of W (n x k) via the eigen-decomposition of its
    Gram matrix; directions with relative Gram eigenvalue < drop_tol^2 are dropped
    (zeroed).  Returns (W, active mask)."""
    G = W.T @ W
    d = np.sqrt(np.maximum(np.diag(G), 1e-300))
    Gs = G / np.outer(d, d)
    lam, Q = np.linalg.eigh(Gs)
    keep = lam > drop_tol * max(lam.max(), 1e-300)
    T = (Q / d[:, None]) * np.where(keep, 1.0 / np.sqrt(np.where(keep, lam, 1.0)), 0.0)[None, :]
    return W @ T, keep


def lobpcg3(apply_L, prec, X0, tol, anorm, maxit=200, AX0=None):
    """Width-3 LOBPCG for the 3 algebraically smallest eigenpairs of the symmetric
    operator ``apply_L`` (n x 3 block in, block out).  ``prec`` applies the SPD
    block-Jacobi preconditioner.  Returns (X, theta, n_applies, resid history).

    Basis S = [X W P] is kept orthonormal; the Rayleigh-Ritz problem is whitened with
    the Cholesky factor of the basis Gram matrix so drift cannot accumulate."""
    n = X0.shape[0]
    napply = 0
    if AX0 is None:
        AX0 = apply_L(X0)
        napply += 1
    # orthonormalise X0 by a linear map (A X C = (A X) C)
    G = X0.T @ X0
    Lc = np.linalg.cholesky(G)
    Ci = np.linalg.inv(Lc).T
    X, AX = X0 @ Ci, AX0 @ Ci
    th, Q = np.linalg.eigh(0.5 * (X.T @ AX + AX.T @ X))
    X, AX = X @ Q, AX @ Q
    P = np.zeros((n, 3))
    AP = np.zeros((n, 3))
    actP = np.zeros(3, bool)
    hist = []
    for it in range(maxit):
        R = AX - X * th[None, :]
        rn = np.linalg.norm(R, axis=0)
        hist.append(rn.max())
        if rn.max() <= tol * anorm:
            break
        W = prec(R)
        Q6 = np.concatenate([X, P], axis=1)
        for _ in range(2):
            W = W - Q6 @ (Q6.T @ W)
        W, actW = _svqb(W)
        for _ in range(1):
            W = W - Q6 @ (Q6.T @ W)
            W, actW2 = _svqb(W)
            actW = actW2
        AW = apply_L(W)
        napply += 1
        S = np.concatenate([X, W, P], axis=1)
        AS = np.concatenate([AX, AW, AP], axis=1)
        act = np.concatenate([np.ones(3, bool), actW, actP])
        Gm = S.T @ AS
        Gm = 0.5 * (Gm + Gm.T)
        Mm = S.T @ S
        for j in range(9):
            if not act[j]:
                Gm[j, :] = 0.0
                Gm[:, j] = 0.0
                Mm[j, :] = 0.0
                Mm[:, j] = 0.0
                Mm[j, j] = 1.0
                Gm[j, j] = BIG
        Rc = np.linalg.cholesky(Mm).T            # Mm = Rc^T Rc
        Ri = np.linalg.inv(Rc)
        lam, Cq = np.linalg.eigh(Ri.T @ Gm @ Ri)
        C = Ri @ Cq[:, :3]                        # 3 smallest (inactive sit at BIG)
        th = lam[:3]
        # new search direction: part of the update outside X, M-orthogonalised against C
        Z = C.copy()
        Z[:3, :] = 0.0
        Z = Z - C @ (C.T @ (Mm @ Z))
        # M-orthonormalise Z (9-dim Gram-Schmidt with drop)
        Cp = np.zeros((9, 3))
        actP = np.zeros(3, bool)
        for j in range(3):
            z = Z[:, j].copy()
            n0 = np.sqrt(max(z @ (Mm @ z), 0.0))
            for i in range(j):
                if actP[i]:
                    z -= Cp[:, i] * (Cp[:, i] @ (Mm @ z))
            for i in range(j):
                if actP[i]:
                    z -= Cp[:, i] * (Cp[:, i] @ (Mm @ z))
            nz = np.sqrt(max(z @ (Mm @ z), 0.0))
            if nz > 1e-8 * max(n0, 1e-300) and nz > 1e-150:
                Cp[:, j] = z / nz
                actP[j] = True
        X, AX = S @ C, AS @ C
        P, AP = S @ Cp, AS @ Cp
    return X, th, napply, hist


def so3sync_model(pc, pt, B, a, n_c, n_t, maxiter, tol=1e-11, stats=None, shortcut=True, spanning_start=True,
                  tol_early=0.0, early_margin=4):
    """Device algorithm for bipgo.py:243-348 (matrix-free, LOBPCG).  Same return
    convention as ``vican_oracle.so3sync``.

    ``spanning_start``: the first eigen-solve starts from project_SO3((P Lambda_T P^T E_0)_c), the one-hop
    estimate around the gauge camera, instead of identity blocks (csrc/rotation.cuh: init_from_root_kernel).
    ``shortcut``: when an eigen-solve accepts its start block R (the previous r_c) without a single step, the
    eigenvectors are V = R C, so V_c V_0^-1 = R_c R_0^T is already a rotation and the primal multiply
    P Lambda_T P^T r_c equals Y R_0^T with Y = P Lambda_T P^T R computed for the eigen-residual: no edge passes
    (csrc/rotation.cuh: so3sync_run).  ``stats.passes`` counts (time, camera) passes per outer iteration.
    ``tol_early`` > 0: inexact inner solves -- outer iterations followed by at least ``early_margin`` more stop their
    eigen-solve at ``tol_early``; the last ``early_margin`` use ``tol`` (vb_so3_options.tol_early).
    ``stats.applies[k] == 1`` means iteration k accepted its start block at the first step."""
    deg_t = np.zeros(n_t)
    np.add.at(deg_t, pt, a)
    deg_c = np.zeros(n_c)
    np.add.at(deg_c, pc, a)
    I3 = np.eye(3)
    LamT = I3[None] / deg_t[:, None, None]
    LamC = I3[None] * deg_c[:, None, None]
    LamCinv = I3[None] / deg_c[:, None, None]
    r_c = None
    r_t = None
    Wt_next = None        # Lambda_T P^T r_c emitted by the dual update: the time half of the next L-apply
    for outer in range(maxiter):
        npass = [0, 0]

        def ppwr(X):      # P Lambda_T P^T X: one time pass + one camera pass
            npass[0] += 1
            npass[1] += 1
            return pass_cam(pc, pt, B, LamT @ pass_time(pc, pt, B, X, n_t), n_c)

        def apply_L(Xf):
            X = Xf.reshape(n_c, 3, 3)
            return (LamC @ X - ppwr(X)).reshape(3 * n_c, 3)

        def prec(Rf):
            return (LamCinv @ Rf.reshape(n_c, 3, 3)).reshape(3 * n_c, 3)

        anorm = 2.0 * np.abs(LamC).sum(axis=(1, 2)).max()
        if r_c is None:
            if spanning_start:
                E0 = np.zeros((n_c, 3, 3))
                E0[0] = I3
                Y0 = ppwr(E0)
                X0b = np.tile(I3, (n_c, 1, 1))
                seen = np.abs(Y0).sum(axis=(1, 2)) > 1e-200
                X0b[seen] = svd_factors(Y0[seen])[0]
            else:
                X0b = np.tile(I3, (n_c, 1, 1))
            Y = ppwr(X0b)
        else:
            X0b = r_c
            npass[1] += 1                                   # camera pass only: the dual update emitted Wt
            Y = pass_cam(pc, pt, B, Wt_next, n_c)
        AX0 = (LamC @ X0b - Y).reshape(3 * n_c, 3)
        tol_k = tol_early if (tol_early > 0.0 and maxiter - outer > early_margin) else tol
        V, th, napp, hist = lobpcg3(apply_L, prec, X0b.reshape(3 * n_c, 3), tol_k, anorm, AX0=AX0)
        if stats is not None:
            stats.applies.append(napp + 1)
            stats.resid.append(hist[-1] / anorm)
            stats.theta.append(th.copy())
        if shortcut and r_c is not None and napp == 0:
            M = Y @ X0b[0].T                                # (P Lambda_T P^T R) R_0^T
        else:
            X = V @ np.linalg.inv(V[:3, :3])
            r_gauge, _, _ = svd_factors(X.reshape(n_c, 3, 3))
            M = ppwr(r_gauge)
        r_c, LamC, LamCinv = svd_factors(M)
        npass[0] += 1
        Yt = pass_time(pc, pt, B, r_c, n_t)
        r_t, _, LamT = svd_factors(Yt)
        Wt_next = LamT @ Yt
        if stats is not None:
            if not hasattr(stats, "passes"):
                stats.passes = []
            stats.passes.append(tuple(npass))
    return r_c, r_t
