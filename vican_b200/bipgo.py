"""Drop-in replacements for the reference's solver entry points.

    from vican_b200.bipgo import bipartite_se3sync, object_bipartite_se3sync

keep the signatures and result conventions of ``vican/bipgo.py:353-490`` and ``:493-545``
(``src_edges`` dict, ``constraints``, ``noise_model_r`` / ``noise_model_t`` / ``edge_filter``
callables, ``maxiter``, ``lsqr_solver``, ``dtype``) and run the numerics on the GPU through
the C ABI (``include/vican_b200.h``).  The Python callables necessarily run on the host, once
per detection, while the dictionary is flattened into arrays; nothing else does.
"""
from __future__ import annotations

import gc
import os
import time as _time
import warnings
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import solver as _solver
from .geometry import SE3

try:                                                                         # built by _cabi.build() / build_flatten()
    from . import _vb_flatten
except ImportError:                                                          # host logic only: the list-comprehension
    _vb_flatten = None                                                       # flatten below is the same statement


def _as_f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)

__all__ = ["bipartite_se3sync", "object_bipartite_se3sync", "large_bipartite_so3sync", "EdgeTable", "solve_table",
           "last_info", "EigenConvergenceWarning"]

# diagnostics of the most recent call (iteration counts, Ritz values, timings)
last_info: Dict[str, object] = {}


class EdgeTable:
    """Flattened detections: what the reference keeps in Python dicts (bipgo.py:203-221,
    :420-431) as arrays, with node indices in the reference's lexicographic node order."""

    def __init__(self, src_edges: dict, constraints: dict, noise_model_r: Callable, noise_model_t: Callable,
                 edge_filter: Callable):
        # the flatten creates millions of short-lived containers; the cyclic collector would walk
        # the whole detection dictionary again and again (measured: 2x on 2 M detections)
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            self._from_dict(src_edges, constraints, noise_model_r, noise_model_t, edge_filter)
        finally:
            if gc_was_on:
                gc.enable()

    def _from_dict(self, src_edges, constraints, noise_model_r, noise_model_t, edge_filter):
        # One pass over the dictionary in insertion order, like the reference's loops; the callables see exactly
        # the kept detections: edge_filter once per detection, the noise models once per KEPT detection
        # (bipgo.py:204, :212, :423, :449).  The walk itself runs in C (csrc/flatten.c: key coding, pose copies,
        # weights) -- what remains per detection is the caller's own callables.
        if _vb_flatten is None or os.environ.get("VICAN_B200_PY_FLATTEN", "0") == "1":
            return self._from_dict_py(src_edges, constraints, noise_model_r, noise_model_t, edge_filter)
        n = len(src_edges)
        R = np.empty((n, 9), dtype=np.float64)
        t = np.empty((n, 3), dtype=np.float64)
        kr = np.empty(n, dtype=np.float64)
        kt = np.empty(n, dtype=np.float64)
        cam_code = np.empty(n, dtype=np.int32)
        tm_code = np.empty(n, dtype=np.int32)
        k, cam_keys, tm_keys, r_fmt, kr0 = _vb_flatten.flatten(src_edges, edge_filter, noise_model_r, noise_model_t,
                                                               _as_f64, R, t, kr, kt, cam_code, tm_code)
        if k == 0:
            raise ValueError("no edge passes edge_filter")
        self._assemble_coded(cam_keys, cam_code[:k], tm_keys, tm_code[:k], R[:k], t[:k], kr[:k], kt[:k], constraints,
                             round_kr_f32=bool(r_fmt == "f" and not isinstance(kr0, np.floating)))

    def _from_dict_py(self, src_edges, constraints, noise_model_r, noise_model_t, edge_filter):
        # the same flatten as list comprehensions (kept as the statement the C walk is tested against)
        kept = [kv for kv in src_edges.items() if edge_filter(kv[1])]
        if not kept:
            raise ValueError("no edge passes edge_filter")
        cams = [kv[0][0] for kv in kept]
        tm_keys = [kv[0][1] for kv in kept]
        Rs = [kv[1]["pose"].R() for kv in kept]
        ts = [kv[1]["pose"].t() for kv in kept]
        kr = [noise_model_r(kv[1]) for kv in kept]                           # bipgo.py:212
        kt = [noise_model_t(kv[1]) for kv in kept]                           # bipgo.py:449
        self._assemble(cams, tm_keys, Rs, ts, kr, kt, constraints)

    @classmethod
    def from_arrays(cls, cam_ids, tm_keys, R, t, k_r, k_t, constraints: dict) -> "EdgeTable":
        """Pre-evaluated fast path: detections that already passed ``edge_filter``
        as parallel sequences -- camera id and ``"<timestamp>_<marker>"`` strings as in the
        reference's keys (cam.py:180), poses as ``[n,3,3]`` / ``[n,3]`` arrays, weights
        ``k_r = noise_model_r(e)``, ``k_t = noise_model_t(e)`` as arrays.  No Python callable runs."""
        self = cls.__new__(cls)
        R = np.asarray(R)
        self._assemble(list(cam_ids), list(tm_keys), R.reshape(-1, 3, 3), np.asarray(t).reshape(-1, 3),
                       np.asarray(k_r, dtype=np.float64), np.asarray(k_t, dtype=np.float64), constraints)
        return self

    def _assemble(self, cams, tm_keys, Rs, ts, kr, kt, constraints):
        n = len(cams)
        if n == 0:
            raise ValueError("no edge passes edge_filter")
        if isinstance(Rs, list) and len({r.dtype for r in Rs}) > 1:
            # the float32 rounding of `k_r * pose.R()` is mirrored per CALL, not per detection
            raise ValueError("detections mix float32 and float64 pose arrays; convert them to one dtype")
        R = np.array(Rs) if isinstance(Rs, list) else Rs                     # np.array: 2.5x faster than np.stack here
        # detections are coded through dictionaries of the DISTINCT key components (first-seen order)
        cam_keys = list(dict.fromkeys(cams))
        tm_dist = list(dict.fromkeys(tm_keys))
        cpos = {c: i for i, c in enumerate(cam_keys)}
        tmpos = {s: i for i, s in enumerate(tm_dist)}
        cam_code = np.fromiter(map(cpos.__getitem__, cams), dtype=np.int32, count=n)
        tm_code = np.fromiter(map(tmpos.__getitem__, tm_keys), dtype=np.int32, count=n)
        self._assemble_coded(cam_keys, cam_code, tm_dist, tm_code, R.astype(np.float64).reshape(-1, 9),
                             (np.array(ts) if isinstance(ts, list) else ts).astype(np.float64).reshape(-1, 3),
                             np.asarray(kr, dtype=np.float64), np.asarray(kt, dtype=np.float64), constraints,
                             round_kr_f32=bool(R.dtype == np.float32 and not isinstance(kr[0], np.floating)))

    def _assemble_coded(self, cam_keys, cam_code, tm_keys, tm_code, R, t, kr, kt, constraints, round_kr_f32):
        """cam_keys / tm_keys: DISTINCT camera ids and "<timestamp>_<marker>" strings; cam_code / tm_code: one index
        into them per kept detection; R [n,9], t [n,3], kr, kt float64."""
        self.root = str(min(list(constraints.keys())))                       # bipgo.py:411 (string min)
        self.n_raw = int(cam_code.shape[0])
        # "timestamp_marker" strings repeat once per observing camera: split each distinct one once
        split = []
        for s in tm_keys:
            parts = s.split("_")                                             # bipgo.py:206-207
            constraints[parts[1]]                                            # KeyError like bipgo.py:209
            split.append((parts[0], parts[1]))
        # Camera order = np.unique over 'c'+id strings (bipgo.py:225-229; the one-letter prefix does not
        # change the order): index 0 is the gauge camera.  Time nodes are labelled in the order of the
        # TRANSLATION unknowns, np.unique over t+'_0' (bipgo.py:420-430), which differs from the
        # rotation stage's np.unique over 't'+t when one timestamp is a prefix of another ('1_0' >
        # '10_0' but 't1' < 't10').  The rotation stage does not depend on how time nodes are
        # labelled; the replayed CSR product of the translation CG does (its row sums run over
        # ascending unknown index, csrc/cg.cuh), so node indices ascend with the unknown index.
        # np.unique runs on the DISTINCT ids only.
        self.cam_ids = np.unique(np.asarray(cam_keys))
        tids = np.unique(np.asarray([t_ + "_0" for t_ in dict.fromkeys(p[0] for p in split)]))
        self.time_ids = np.asarray([t_[:-2] for t_ in tids])
        self.marker_ids = sorted({p[1] for p in split} | {self.root})
        cpos = {str(c): i for i, c in enumerate(self.cam_ids)}
        tpos = {str(t_): i for i, t_ in enumerate(self.time_ids)}
        mpos = {m: i for i, m in enumerate(self.marker_ids)}
        cam_remap = np.fromiter((cpos[c] for c in cam_keys), dtype=np.int32, count=len(cam_keys))
        t_remap = np.fromiter((tpos[p[0]] for p in split), dtype=np.int32, count=len(split))
        m_remap = np.fromiter((mpos[p[1]] for p in split), dtype=np.int32, count=len(split))
        self.cam_idx = cam_remap[cam_code]
        self.time_idx = t_remap[tm_code]
        self.marker_idx = m_remap[tm_code]
        # numpy evaluates `k_r * pose.R()` in float32 when the pose arrays are float32 (poses
        # that went through SE3.inv(), geometry.py:209-211) and k_r is a Python float.  Only that first
        # product is rounded here; the reference's float32 chain also rounds the two constraint products
        # when the constraints are float32 arrays (deviation <= 1e-7 rad, tests/golden f32 case).
        self.round_kr_f32 = bool(round_kr_f32)
        self.R = R
        self.t = t
        self.k_r = kr
        self.k_t = kt
        # per-marker constants (<= a few dozen): same numpy expressions as the reference
        R0 = np.asarray(constraints[self.root].R())
        self.markerC = np.stack([np.asarray(constraints[m].R(), dtype=np.float64).T @ R0.astype(np.float64)
                                 for m in self.marker_ids])                  # R_m^T R_0, bipgo.py:213
        q = []
        for m in self.marker_ids:
            r_0m = constraints[self.root].R().T @ constraints[m].R()         # bipgo.py:451
            t_m0 = (constraints[m].inv() @ constraints[self.root]).t()       # bipgo.py:452 (float32 arithmetic)
            q.append(np.asarray(r_0m, dtype=np.float64) @ np.asarray(t_m0, dtype=np.float64))
        self.marker_q = np.stack(q)

    def unknown_index(self):
        """(unk_c, unk_t): position of every camera / time node in the reference's unknown vector,
        np.unique over camera ids and t+'_0' strings together (bipgo.py:420-430)."""
        names = np.asarray(list(self.cam_ids) + [t + "_0" for t in self.time_ids])
        order = np.argsort(names, kind="stable")
        unk = np.empty(order.shape[0], dtype=np.int32)
        unk[order] = np.arange(order.shape[0], dtype=np.int32)
        return unk[:self.n_c], unk[self.n_c:]

    @property
    def n_c(self) -> int:
        return int(self.cam_ids.shape[0])

    @property
    def n_t(self) -> int:
        return int(self.time_ids.shape[0])


class EigenConvergenceWarning(RuntimeWarning):
    """The eigen-iteration of an outer iteration stopped at its step cap before reaching its
    tolerance (the reference's ARPACK call raises ArpackNoConvergence in that situation)."""


def _solve_table(tab: EdgeTable, maxiter: int, lsqr_solver: Optional[str], mode: str = "parity",
                 tol: float = 1e-13, strict: bool = False, verbose: bool = False):
    t0 = _time.perf_counter()
    if tab.n_c < 3:
        raise ValueError("the rotation stage needs at least 3 camera nodes (the reference asks ARPACK for 5 "
                         "eigenpairs of a 3 n_c x 3 n_c matrix); got %d" % tab.n_c)
    g = _solver.DeviceGraph(tab.cam_idx, tab.time_idx, tab.marker_idx, tab.R, tab.k_r, tab.k_t, tab.markerC,
                            tab.n_c, tab.n_t, round_kr_f32=tab.round_kr_f32)
    # The reference leaves its loop early when the five eigenvalues nearest zero are all <= 1e-6
    # (bipgo.py:283-292).  lambda_4, lambda_5 are O(degree) on a connected graph, so the test can only fire when
    # the graph falls apart into components: only then (or for verbose callers, who get the reference's
    # evals / eigengap read-out) is the second eigen-solve per iteration paid for.
    n_comp = g.n_components()
    if n_comp > 1:
        warnings.warn("the detection graph has %d connected components: poses outside the gauge camera's component "
                      "are undetermined (the reference returns an arbitrary member of a 3 x %d dimensional "
                      "eigenspace there)" % (n_comp, n_comp), RuntimeWarning, stacklevel=3)
    rot = _solver.solve_rotations(g, maxiter, tol=tol, eval_gap=bool(verbose or n_comp > 1))
    if verbose:
        for it in range(min(rot.stats.outer_done, 64)):
            ev = list(rot.stats.evals_hist[it])
            gap = abs(ev[3] / ev[2]) if ev[2] != 0.0 else float("inf")       # bipgo.py:291
            print("Optimizing %d/%d: evals0=%1.3e, evals1=%1.3e, evals2=%1.3e, eigengap=%1.3e"
                  % (it + 1, maxiter, ev[0], ev[1], ev[2], gap))
        if rot.stats.early_exit:
            print("stopped after %d of %d iterations: max |eval| <= 1e-6" % (rot.stats.outer_done, maxiter))
    if rot.status == 2 or rot.stats.stalled_outer > 0:
        msg = ("eigen-iteration stopped at its step cap in %d of %d outer iterations (residual %.2e, scale %.2e): "
               "the poses may be inaccurate (outliers / weak connectivity?)"
               % (rot.stats.stalled_outer, maxiter, max(rot.stats.resid), rot.stats.anorm))
        if strict:
            raise _solver.ConvergenceError(msg)
        warnings.warn(msg, EigenConvergenceWarning, stacklevel=3)
    tr = None
    if lsqr_solver is not None:
        tr = _solver.solve_translations(g, rot, tab.t, tab.marker_q, lsqr_solver, mode=mode,
                                        unknown_index=tab.unknown_index())
    torch.cuda.synchronize()
    last_info.clear()
    last_info.update(dict(
        n_c=g.n_c, n_t=g.n_t, n_edges=g.n_edges, n_raw=g.n_raw, n_tiles=g.n_tiles,
        inner_per_outer=list(rot.stats.inner_per_outer[:min(maxiter, 64)]),
        time_passes=rot.stats.time_passes, cam_passes=rot.stats.cam_passes,
        theta=list(rot.stats.theta), resid=list(rot.stats.resid), eig_status=rot.status,
        outer_done=rot.stats.outer_done, early_exit=bool(rot.stats.early_exit), n_components=n_comp,
        evals=[list(rot.stats.evals_hist[i]) for i in range(min(rot.stats.outer_done, 64))] if (verbose or n_comp > 1) else None,
        trans_iters=None if tr is None else tr.iters, trans_istop=None if tr is None else tr.istop,
        device_seconds=_time.perf_counter() - t0))
    return g, rot, tr


def large_bipartite_so3sync(src_edges: dict, constraints: dict, noise_model: Callable, edge_filter: Callable,
                            maxiter: int, dtype=np.float32, *, verbose: bool = False) -> dict:
    """Rotation stage only (vican/bipgo.py:145-350): {camera id: R, f"{t}_0": R} wrt the world.
    ``verbose`` prints the reference's per-iteration read-out (evals0..2, eigengap; bipgo.py:336-339)."""
    tab = EdgeTable(src_edges, constraints, noise_model, lambda e: 1.0, edge_filter)
    _, rot, _ = _solve_table(tab, maxiter, None, verbose=verbose)
    Rc, Rt = rot.world_rotations()
    Rc, Rt = Rc.cpu().numpy().astype(dtype), Rt.cpu().numpy().astype(dtype)
    out = {}
    for i, c in enumerate(tab.cam_ids):                                      # bipgo.py:344-348
        out[c] = Rc[i]
    for j in np.argsort(tab.time_ids, kind="stable"):                        # insertion order: np.unique('t' + t)
        out[tab.time_ids[j] + "_0"] = Rt[j]
    return out


def bipartite_se3sync(src_edges: dict, constraints: dict, noise_model_r: Callable, noise_model_t: Callable,
                      edge_filter: Callable, maxiter: int, lsqr_solver: str, dtype=np.float32,
                      *, mode: str = "parity", strict: bool = False, verbose: bool = False) -> dict:
    """SE(3) synchronisation in a bipartite camera / object-timestep graph with node
    constraints; same contract as ``vican/bipgo.py:353-490``.  Returns a dict with every camera
    id and every ``f"{t}_0"`` node mapped to an ``SE3`` pose wrt the world.

    The arithmetic is always fp64 on the device; ``dtype`` only selects the dtype of the returned
    rotation arrays (the reference's ``R`` follows ``dtype``, its ``t`` is float64).
    Raises ``ValueError`` for an unknown ``lsqr_solver`` (the reference falls through to a
    NameError) and ``ConvergenceError`` (an ``AssertionError``) if CG does not converge.  An
    eigen-iteration that stops at its step cap emits ``EigenConvergenceWarning`` (``strict=True``:
    raises ``ConvergenceError``, like ARPACK's ``ArpackNoConvergence`` in the reference).
    ``verbose=True`` prints what the reference's progress bar shows (evals0..2, eigengap per iteration,
    bipgo.py:336-339).  The reference's early exit ``max |lambda_1..5| <= 1e-6`` (bipgo.py:283) is applied
    whenever it can fire, i.e. when the detection graph is not connected (checked on the host)."""
    if lsqr_solver not in ("conjugate_gradient", "direct"):
        raise ValueError("lsqr_solver must be 'conjugate_gradient' or 'direct', got %r" % (lsqr_solver,))
    tab = EdgeTable(src_edges, constraints, noise_model_r, noise_model_t, edge_filter)
    return solve_table(tab, maxiter, lsqr_solver, dtype=dtype, mode=mode, strict=strict, verbose=verbose)


def solve_table(tab: EdgeTable, maxiter: int, lsqr_solver: str, dtype=np.float32, *, mode: str = "parity",
                strict: bool = False, verbose: bool = False) -> dict:
    """``bipartite_se3sync`` from an already flattened ``EdgeTable`` (``EdgeTable.from_arrays``,
    ``io.EdgeAccumulator.table``): same result dictionary, no dictionary walk, no callables."""
    if lsqr_solver not in ("conjugate_gradient", "direct"):
        raise ValueError("lsqr_solver must be 'conjugate_gradient' or 'direct', got %r" % (lsqr_solver,))
    _, rot, tr = _solve_table(tab, maxiter, lsqr_solver, mode=mode, strict=strict, verbose=verbose)
    Rc, Rt = rot.world_rotations()
    Rc, Rt = Rc.cpu().numpy().astype(dtype), Rt.cpu().numpy().astype(dtype)
    xc, xt = tr.x_c.cpu().numpy(), tr.x_t.cpu().numpy()
    # result order = np.unique over camera ids and t+'_0' (bipgo.py:420-430, :484-487)
    cam_keys = list(tab.cam_ids)
    time_keys = [t + "_0" for t in tab.time_ids]
    names = np.asarray(cam_keys + time_keys)
    order = np.argsort(names, kind="stable")
    Rall = np.concatenate([Rc, Rt], axis=0)
    tall = np.concatenate([xc, xt], axis=0)
    out = {}
    for i in order:
        out[names[i]] = SE3(R=Rall[i], t=tall[i])
    return out


def object_bipartite_se3sync(src_edges: dict, noise_model_r: Callable, noise_model_t: Callable,
                             edge_filter: Callable, maxiter: int, lsqr_solver: str, dtype=np.float32,
                             *, mode: str = "parity", strict: bool = False, verbose: bool = False) -> dict:
    """Object calibration (single object, moving camera); same contract as
    ``vican/bipgo.py:493-545``: markers take the camera role, timesteps the object role, every
    pose is inverted (through float32, as ``SE3.inv`` does) and only marker poses are returned."""
    root = str(min([int(k[1].split("_")[1]) for k in src_edges.keys()]))      # bipgo.py:524 (numeric min)
    # bipgo.py:526-531: re-key and invert every pose.  SE3.inv() stores its result in a float32
    # 4x4 (geometry.py:239-243); fp64 poses are inverted as one device batch with the same
    # rounding, float32 poses keep numpy's float32 arithmetic via the container itself.
    gc_was_on = gc.isenabled()
    gc.disable()                       # tens of thousands of new containers: keep the cyclic collector off the big dict
    try:
        edges = _rekey_inverted(src_edges, root)
    finally:
        if gc_was_on:
            gc.enable()
    out = bipartite_se3sync(edges, constraints={root: SE3(pose=np.eye(4))}, noise_model_r=noise_model_r,
                            noise_model_t=noise_model_t, edge_filter=edge_filter, maxiter=maxiter,
                            lsqr_solver=lsqr_solver, dtype=dtype, mode=mode, strict=strict, verbose=verbose)
    return {k: v for k, v in out.items() if "_" not in k}                    # bipgo.py:543


def _rekey_inverted(src_edges: dict, root: str) -> dict:
    """bipgo.py:526-531: {(marker, "t_root"): {pose: inverted pose, corners, reprojected_err, im_filename}}."""
    keys = list(src_edges.keys())
    vals = list(src_edges.values())
    use_c = _vb_flatten is not None and os.environ.get("VICAN_B200_PY_FLATTEN", "0") != "1"
    R0 = np.asarray(vals[0]["pose"].R()) if vals else None
    if vals and R0.dtype == np.float64:
        from . import ops
        if use_c:                                                            # pose arrays stacked in C (csrc/flatten.c)
            Rs = np.empty((len(vals), 3, 3), dtype=np.float64)
            ts = np.empty((len(vals), 3), dtype=np.float64)
            _vb_flatten.poses(vals, _as_f64, Rs, ts)
        else:
            Rs = np.stack([v["pose"].R() for v in vals])
            ts = np.stack([np.asarray(v["pose"].t(), dtype=np.float64) for v in vals])
        Ri, ti = ops.se3_invert_batch(Rs, ts, round_f32=True)
        P4 = np.zeros((len(vals), 4, 4), dtype=np.float32)
        P4[:, :3, :3] = Ri.cpu().numpy()
        P4[:, :3, 3] = ti.cpu().numpy()
        P4[:, 3, 3] = 1.0
        inv_poses = list(map(SE3._from_pose32, P4))
    else:
        inv_poses = [v["pose"].inv() for v in vals]
    if use_c:
        edges = _vb_flatten.rekey(keys, vals, inv_poses, root)
    else:
        edges = {}
        for k, v, ip in zip(keys, vals, inv_poses):
            t, marker_id = k[1].split("_")
            edges[marker_id, t + "_" + root] = {"pose": ip,
                                                "corners": v["corners"],
                                                "reprojected_err": v["reprojected_err"],
                                                "im_filename": v["im_filename"]}
    return edges
