"""CPU ORACLE -- test infrastructure, NOT product code.

Vectorised numpy/scipy restatement of the reference's bipartite SE(3)
synchronisation (``/root/reference/vican/bipgo.py``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this module; nothing under ``vican_b200/`` does.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md 8c), so the oracle is pinned against OUTPUTS OF THE REAL REFERENCE run
in the build container: ``tests/golden/*.npz`` were produced by
``tests/golden/make_golden.py`` (imports ``/root/reference/vican``) and
``tests/test_oracle_vs_golden.py`` checks this restatement against them.

Third-party arithmetic (not under /root/reference; the reference pins
scipy==1.5.4 / numpy==1.19.5, the container has scipy 1.18 / numpy 2.3): the
oracle calls the same scipy routines the reference calls (``cg``, ``lsqr``) on the
same matrices, and replaces ARPACK ``eigs`` by a dense symmetric ``eigh`` with the
reference's selection rule (3 eigenvalues nearest sigma=-1e-6); SURVEY.md 4.2
measured this equivalent to <= 6e-15 rad.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import cg as _scipy_cg, lsqr as _scipy_lsqr


# --------------------------------------------------------------------------- utils
def svd_polar_batch(M: np.ndarray):
    """Per-block SVD factors used by bipgo.py:306-312 / :323-329 and
    geometry.py:189-190.  Returns (rot, U, S) with rot = U diag(1,1,det(U Vt)) Vt."""
    U, S, Vt = np.linalg.svd(M)
    d = np.linalg.det(U @ Vt)
    D = np.zeros_like(M)
    D[:, 0, 0] = 1.0
    D[:, 1, 1] = 1.0
    D[:, 2, 2] = d
    return U @ D @ Vt, U, S


def _block_matrix(rows, cols, blocks, n_rows, n_cols):
    """Sparse (3 n_rows x 3 n_cols) matrix of 3x3 blocks, bipgo.py:252-270 layout."""
    E = rows.shape[0]
    ii = (3 * rows[:, None, None] + np.arange(3)[None, :, None] + np.zeros((1, 1, 3), dtype=np.int64))
    jj = (3 * cols[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 3, 1), dtype=np.int64))
    return sp.csr_matrix((blocks.reshape(-1), (ii.reshape(-1), jj.reshape(-1))),
                         shape=(3 * n_rows, 3 * n_cols))


# --------------------------------------------------------------- rotation stage
def fold_blocks(R, k_r, marker, marker_R, root):
    """bipgo.py:209-213: blk = (k_r R_cm) R_m^T R_0, evaluated left to right.  When the
    pose arrays are float32 (object variant: ``pose.inv()`` returns float32 views,
    geometry.py:209-211) and k_r is a Python float, numpy evaluates ``k_r * R`` in
    float32; ``k_r`` given as a float32 array reproduces that rounding."""
    kR = k_r[:, None, None] * R
    return (kR.astype(np.float64) @ np.transpose(marker_R, (0, 2, 1))[marker]) @ marker_R[root]


def aggregate_pairs(cam, time, blk, k_r, n_t):
    """bipgo.py:215-221: sum blocks and weights per (camera, time) pair; pairs are
    returned in first-occurrence (dict insertion) order."""
    key = cam.astype(np.int64) * n_t + time
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    pid = rank[inv]
    P = uniq.shape[0]
    B = np.zeros((P, 3, 3))
    np.add.at(B, pid, blk)
    a = np.zeros(P)
    np.add.at(a, pid, k_r.astype(np.float64))
    return cam[first[order]], time[first[order]], B, a


def so3sync(pair_cam, pair_time, B, a, n_c, n_t, maxiter, sigma=-1e-6, return_history=False, timings=None):
    """bipgo.py:243-348 on aggregated edges.  Returns (r_c, r_t) as stored by
    the reference BEFORE the final transpose (bipgo.py:344-348), i.e. the world
    rotations are r_c[i].T and r_t[j].T.  ``timings`` (a list) receives the wall time of every
    primal-dual iteration (bench.py's CPU arm)."""
    import time as _time
    P = _block_matrix(pair_cam, pair_time, B, n_c, n_t)                               # :269
    adj = sp.csr_matrix((a, (pair_cam, pair_time)), shape=(n_c, n_t))                 # :270
    deg_t = np.asarray(adj.sum(axis=0)).ravel()                                       # :271
    Ppwr = P @ sp.diags(1.0 / np.repeat(deg_t, 3)) @ P.T                              # :273
    pwr_adj = adj @ sp.diags(1.0 / deg_t) @ adj.T                                     # :274
    pwr_deg = np.asarray(pwr_adj.sum(axis=-1)).ravel()                                # :275
    lbd_c = sp.diags(np.repeat(pwr_deg, 3))                                           # :276
    hist = []
    r_c = r_t = None
    max_eval = 1.0
    bidx_c = np.arange(n_c)
    bidx_t = np.arange(n_t)
    for _ in range(maxiter):                                                          # :282
        if max_eval <= 1e-6:                                                          # :283
            break
        _t0 = _time.perf_counter()
        L = (lbd_c - Ppwr)
        L = 0.5 * (L + L.T)                                                           # :285-286
        evals, evecs = np.linalg.eigh(L.toarray())
        sel = np.argsort(np.abs(evals - sigma), kind="stable")[:5]                    # :288 (k=5 nearest sigma)
        ev5 = evals[sel]
        max_eval = np.abs(ev5).max()                                                  # :292
        V = evecs[:, sel[:3]]
        hist.append(ev5.copy())
        X = V @ np.linalg.inv(V[:3, :3])                                              # :295
        r_c, _, _ = svd_polar_batch(X.reshape(n_c, 3, 3))                             # :296-297
        M = (Ppwr @ r_c.reshape(3 * n_c, 3)).reshape(n_c, 3, 3)                       # :300
        r_c, U, S = svd_polar_batch(M)                                                # :306-308
        lc = (U * S[:, None, :]) @ np.transpose(U, (0, 2, 1))                         # :312
        lbd_c = _block_matrix(bidx_c, bidx_c, lc, n_c, n_c)
        Y = (P.T @ r_c.reshape(3 * n_c, 3)).reshape(n_t, 3, 3)                        # :318
        r_t, U, S = svd_polar_batch(Y)                                                # :323-325
        lt = (U * (1.0 / S)[:, None, :]) @ np.transpose(U, (0, 2, 1))                 # :329
        lbd_t = _block_matrix(bidx_t, bidx_t, lt, n_t, n_t)
        Ppwr = P @ lbd_t @ P.T                                                        # :334
        if timings is not None:
            timings.append(_time.perf_counter() - _t0)
    if return_history:
        return r_c, r_t, hist
    return r_c, r_t


# ------------------------------------------------------------ translation stage
def translation_system(cam, time, marker, t_cm, k_t, marker_R, marker_t_inv0, root,
                       Rw_c, Rw_t, n_c, n_t, node_cam, node_time):
    """bipgo.py:445-471.  ``Rw_c``/``Rw_t`` are the world rotations returned by the
    rotation stage (already transposed); ``node_cam``/``node_time`` map camera /
    time indices to the row of the unknown vector (reference node order, :429-430);
    ``marker_t_inv0[m]`` = (T_m^-1 T_0).t()  (:452, float32 arithmetic in the
    reference -- computed by the caller with the reference's own container).
    Returns (J, t_tilde)."""
    E = cam.shape[0]
    N = n_c + n_t
    r_0m = np.transpose(marker_R[root])[None] @ marker_R                              # :451
    q = np.einsum("eij,ej->ei", Rw_t[time] @ r_0m[marker], marker_t_inv0[marker])
    d = np.einsum("eij,ej->ei", Rw_c[cam], t_cm) + q                                  # :454-455
    t_tilde = (k_t[:, None] * d).reshape(-1)                                          # :461
    rows = (3 * np.arange(E)[:, None] + np.arange(3)[None]).reshape(-1)
    cols_c = (3 * node_cam[cam][:, None] + np.arange(3)[None]).reshape(-1)
    cols_t = (3 * node_time[time][:, None] + np.arange(3)[None]).reshape(-1)
    vals = np.repeat(k_t, 3)
    J = sp.csr_matrix((np.concatenate([-vals, vals]),
                       (np.concatenate([rows, rows]), np.concatenate([cols_c, cols_t]))),
                      shape=(3 * E, 3 * N))                                           # :463-471
    return J, t_tilde


def solve_translations(J, t_tilde, lsqr_solver):
    """bipgo.py:476-480 with the container's scipy (cg rtol=1e-5; lsqr atol=btol=1e-6)."""
    if lsqr_solver == "conjugate_gradient":
        x, code = _scipy_cg(J.T @ J, J.T @ t_tilde)
        assert code == 0
        return x, dict(code=code)
    if lsqr_solver == "direct":
        x, istop, itn, normr = _scipy_lsqr(J, t_tilde)[:4]
        return x, dict(istop=istop, itn=itn, normr=normr)
    raise ValueError(lsqr_solver)


# ------------------------------------------------------------------ dict front-end
def _flatten(src_edges, edge_filter):
    """Filter + parse keys, dict-insertion order (bipgo.py:203-207, :422-427)."""
    cams, times, marks, Rs, ts, vals = [], [], [], [], [], []
    for e, v in src_edges.items():
        if edge_filter(v):
            tstr, mstr = e[1].split("_")
            cams.append(e[0])
            times.append(tstr)
            marks.append(mstr)
            Rs.append(v["pose"].R())
            ts.append(v["pose"].t())
            vals.append(v)
    return cams, times, marks, Rs, ts, vals


def bipartite_se3sync_oracle(src_edges, constraints, noise_model_r, noise_model_t, edge_filter,
                             maxiter, lsqr_solver, return_info=False):
    """Restatement of bipgo.py:353-490 (dtype=np.float64 behaviour).  Returns
    {node id: (R, t)} for every camera id and every ``f"{t}_0"`` node."""
    root = str(min(list(constraints.keys())))                                         # :411
    cams, times, marks, Rs, ts, vals = _flatten(src_edges, edge_filter)
    # node order: np.unique over 'c'+id / 't'+timestamp strings (:225-229)
    ucam, cam_idx = np.unique(np.array(["c" + c for c in cams]), return_inverse=True)
    utime, time_idx = np.unique(np.array(["t" + t for t in times]), return_inverse=True)
    umark = sorted(set(marks) | {root})
    mpos = {m: i for i, m in enumerate(umark)}
    mark_idx = np.array([mpos[m] for m in marks], dtype=np.int64)
    marker_R = np.stack([np.asarray(constraints[m].R(), dtype=np.float64) for m in umark])
    k_r_list = [noise_model_r(v) for v in vals]                                       # :212
    k_r = np.array(k_r_list, dtype=np.float64)
    R = np.stack(Rs)
    f32_product = (R.dtype == np.float32) and not isinstance(k_r_list[0], np.floating)
    n_c, n_t = ucam.shape[0], utime.shape[0]
    blk = fold_blocks(R if f32_product else R.astype(np.float64),
                      k_r.astype(np.float32) if f32_product else k_r,
                      mark_idx, marker_R, mpos[root])
    pc, pt, B, a = aggregate_pairs(cam_idx, time_idx, blk, k_r, n_t)
    r_c, r_t, hist = so3sync(pc, pt, B, a, n_c, n_t, maxiter, return_history=True)
    Rw_c = np.transpose(r_c, (0, 2, 1))                                               # :344-348
    Rw_t = np.transpose(r_t, (0, 2, 1))
    # translation node order: np.unique over cam ids and t+'_0' (:420-430)
    cam_ids = [s[1:] for s in ucam]
    time_ids = [s[1:] + "_0" for s in utime]
    nodes = np.unique(np.array(cam_ids + time_ids))
    node2idx = {n: i for i, n in enumerate(nodes)}
    node_cam = np.array([node2idx[c] for c in cam_ids])
    node_time = np.array([node2idx[t] for t in time_ids])
    marker_t_inv0 = np.stack([np.asarray((constraints[m].inv() @ constraints[root]).t(), dtype=np.float64)
                              for m in umark])                                        # :452
    k_t = np.array([noise_model_t(v) for v in vals], dtype=np.float64)                # :449
    t_cm = np.stack(ts).astype(np.float64)
    J, t_tilde = translation_system(cam_idx, time_idx, mark_idx, t_cm, k_t, marker_R, marker_t_inv0,
                                    mpos[root], Rw_c, Rw_t, n_c, n_t, node_cam, node_time)
    x, info = solve_translations(J, t_tilde, lsqr_solver)
    x = x.reshape(-1, 3)
    out = {}
    for i, c in enumerate(cam_ids):
        out[c] = (Rw_c[i], x[node_cam[i]])
    for j, t in enumerate(time_ids):
        out[t] = (Rw_t[j], x[node_time[j]])
    if return_info:
        info["evals"] = hist
        info["n_c"], info["n_t"], info["E"], info["E_raw"] = n_c, n_t, pc.shape[0], cam_idx.shape[0]
        return out, info
    return out


def object_bipartite_se3sync_oracle(src_edges, noise_model_r, noise_model_t, edge_filter,
                                    maxiter, lsqr_solver, se3_cls, return_info=False):
    """Restatement of bipgo.py:493-545; ``se3_cls`` supplies the float32-rounding
    container (the reference's or vican_b200.geometry.SE3)."""
    root = str(min(int(k[1].split("_")[1]) for k in src_edges.keys()))                # :524
    edges = {}
    for k, v in src_edges.items():                                                    # :526-531
        t, m = k[1].split("_")
        nv = dict(v)
        nv["pose"] = v["pose"].inv()
        edges[m, t + "_" + root] = nv
    res = bipartite_se3sync_oracle(edges, {root: se3_cls(pose=np.eye(4))}, noise_model_r, noise_model_t,
                                   edge_filter, maxiter, lsqr_solver, return_info=return_info)
    out, info = res if return_info else (res, None)
    out = {k: v for k, v in out.items() if "_" not in k}                              # :543
    return (out, info) if return_info else out


# ------------------------------------------------------------------ array front-end
def solve_arrays_oracle(cam, time, marker, R, t, k_r, k_t, marker_R, marker_t_inv0, root, n_c, n_t, maxiter,
                        lsqr_solver, timings=None):
    """Whole path on pre-indexed arrays (used by bench.py's CPU baseline and the large-shape
    tests): bipgo.py:203-348 + :434-487 without the dictionaries.  Unknown order is
    [cameras; time nodes].  Returns (Rw_c, Rw_t, x_c, x_t)."""
    blk = fold_blocks(R, k_r, marker, marker_R, root)
    pc, pt, B, a = aggregate_pairs(cam, time, blk, k_r, n_t)
    r_c, r_t = so3sync(pc, pt, B, a, n_c, n_t, maxiter, timings=timings)
    Rw_c = np.transpose(r_c, (0, 2, 1))
    Rw_t = np.transpose(r_t, (0, 2, 1))
    J, t_tilde = translation_system(cam, time, marker, t, k_t, marker_R, marker_t_inv0, root, Rw_c, Rw_t, n_c, n_t,
                                    np.arange(n_c), n_c + np.arange(n_t))
    x, _ = solve_translations(J, t_tilde, lsqr_solver)
    x = x.reshape(-1, 3)
    return Rw_c, Rw_t, x[:n_c], x[n_c:]


# ------------------------------------------------------- evaluation helpers (geometry.py)
def angle_deg_oracle(R: np.ndarray) -> np.ndarray:
    """geometry.py:131-151 ``angle``: degrees(arccos(clip((trace(r) - 1) / 2, -1, 1))), batched."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    return np.rad2deg(np.arccos(np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)))


def distance_SO3_oracle(R1: np.ndarray, R2: np.ndarray) -> np.ndarray:
    """geometry.py:154-172 ``distance_SO3``: angle(r1.T @ r2), batched."""
    R1 = np.asarray(R1, dtype=np.float64).reshape(-1, 3, 3)
    R2 = np.asarray(R2, dtype=np.float64).reshape(-1, 3, 3)
    return angle_deg_oracle(np.transpose(R1, (0, 2, 1)) @ R2)


def optimize_gauge_SE3_oracle(Ra, ta, Rb, tb):
    """geometry.py:294-324 ``optimize_gauge_SE3`` on stacked arrays: sum = sum a.R^T b.R
    (:317), gauge_t = sum b.R^T (a.t - b.t) / n (:318, :322), gauge_r = project_SO3(sum^T)
    (:320-321).  ``ta = tb = None`` gives ``optimize_gauge_SO3`` (:264-291)."""
    Ra = np.asarray(Ra, dtype=np.float64).reshape(-1, 3, 3)
    Rb = np.asarray(Rb, dtype=np.float64).reshape(-1, 3, 3)
    s = (np.transpose(Ra, (0, 2, 1)) @ Rb).sum(axis=0)
    u, _, vh = np.linalg.svd(s.T)
    gr = u @ np.diag([1.0, 1.0, np.linalg.det(u @ vh)]) @ vh
    if ta is None:
        return gr, None
    d = np.asarray(ta, dtype=np.float64).reshape(-1, 3) - np.asarray(tb, dtype=np.float64).reshape(-1, 3)
    gt = np.einsum("nji,nj->ni", Rb, d).sum(axis=0) / Ra.shape[0]
    return gr, gt
