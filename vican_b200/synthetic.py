"""Synthetic camera-network graphs (SURVEY.md 8d recipe).

The Google-Drive datasets of the reference are not available offline, so every
test/bench input is generated here.  The generator is array-native (numpy); a
dict exporter builds the ``src_edges`` / ``constraints`` dictionaries in the
format the reference's detection stage produces (``vican/cam.py:59-80,180-185``):
key ``(cam_id, f"{timestep}_{marker_id}")``, value ``{'pose', 'corners',
'reprojected_err', 'im_filename'}`` (+ a weight field ``'w'`` read by the
synthetic noise models).

Conventions (verified against the reference):
detection pose = T_c^-1 . T_t . T_0^-1 . T_m  (marker m at time t in camera c's frame),
``constraints[m]`` = T_m (marker pose in the object frame).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, Optional, Tuple

import numpy as np

__all__ = [
    "SyntheticGraph", "random_rotations", "so3_exp", "make_camera_network",
    "make_object_calibration", "to_edge_dict", "default_callables", "CONFIGS",
]


def random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    """Uniform (Haar) rotations from normalised Gaussian quaternions."""
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - z * w)
    R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w)
    R[:, 2, 1] = 2 * (y * z + x * w)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def so3_exp(xi: np.ndarray) -> np.ndarray:
    """Rodrigues formula for a batch of rotation vectors (n,3) -> (n,3,3)."""
    xi = np.asarray(xi, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(xi, axis=1)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th * th / 6.0, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(ths)) / (ths * ths))
    K = np.zeros((xi.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -xi[:, 2], xi[:, 1]
    K[:, 1, 0], K[:, 1, 2] = xi[:, 2], -xi[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -xi[:, 1], xi[:, 0]
    return np.eye(3)[None] + a[:, None, None] * K + b[:, None, None] * (K @ K)


@dataclasses.dataclass
class SyntheticGraph:
    """Array form of a detection graph plus its ground truth.

    ``cam``/``time``/``marker`` are integer ids (their *string* forms are the
    dictionary ids); ``R``/``t`` the detection poses; ``w`` the per-edge weight
    field; ``reproj`` the per-edge reprojection error (outliers get 1.0).
    """
    n_cams: int
    n_times: int
    n_markers: int
    cam: np.ndarray
    time: np.ndarray
    marker: np.ndarray
    R: np.ndarray
    t: np.ndarray
    w: np.ndarray
    reproj: np.ndarray
    marker_R: np.ndarray          # constraints (marker -> object)
    marker_t: np.ndarray
    gt_cam_R: np.ndarray          # camera -> world
    gt_cam_t: np.ndarray
    gt_obj_R: np.ndarray          # root marker at time t -> world
    gt_obj_t: np.ndarray
    kind: str = "network"         # "network" | "object"

    @property
    def n_edges(self) -> int:
        return int(self.cam.shape[0])


def _visibility(rng, n_cams, n_times, n_markers, cams_per_t, marks_per_cam):
    """Each timestep is seen by ``cams_per_t`` distinct cameras, each seeing
    ``marks_per_cam`` distinct markers (render.py:348-371: >= 2 cameras/timestep)."""
    # argpartition of iid uniforms = uniform sampling without replacement, vectorised
    cu = rng.random((n_times, n_cams))
    cams = np.argpartition(cu, cams_per_t - 1, axis=1)[:, :cams_per_t]
    cams.sort(axis=1)
    mu = rng.random((n_times * cams_per_t, n_markers))
    marks = np.argpartition(mu, marks_per_cam - 1, axis=1)[:, :marks_per_cam]
    marks.sort(axis=1)
    time = np.repeat(np.arange(n_times), cams_per_t * marks_per_cam)
    cam = np.repeat(cams.reshape(-1), marks_per_cam)
    marker = marks.reshape(-1)
    return cam.astype(np.int64), time.astype(np.int64), marker.astype(np.int64)


def _cube_markers(side: float = 0.575) -> Tuple[np.ndarray, np.ndarray]:
    """24-marker cube: 6 faces x 4 markers, side 0.575 m (render.py:468-469)."""
    def rx(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])

    def ry(a):
        c, s = np.cos(a), np.sin(a)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    faces = [np.eye(3), ry(np.pi / 2), ry(np.pi), ry(-np.pi / 2), rx(np.pi / 2), rx(-np.pi / 2)]
    h, q = side / 2, side / 4
    Rs, ts = [], []
    for F in faces:
        for (u, v) in [(-q, -q), (q, -q), (q, q), (-q, q)]:
            Rs.append(F)
            ts.append(F @ np.array([u, v, h]))
    return np.stack(Rs), np.stack(ts)


def make_camera_network(seed: int, n_cams: int, n_times: int, n_markers: int,
                        cams_per_t: int, marks_per_cam: int,
                        sigma_R: float = 0.02, sigma_t: float = 0.01,
                        outlier_frac: float = 0.0, cube: bool = False) -> SyntheticGraph:
    rng = np.random.default_rng(seed)
    cam_R = random_rotations(rng, n_cams)
    cam_t = rng.normal(0.0, 5.0, (n_cams, 3))
    obj_R = random_rotations(rng, n_times)
    obj_t = rng.normal(0.0, 3.0, (n_times, 3))
    if cube and n_markers == 24:
        mk_R, mk_t = _cube_markers()
    else:
        mk_R = random_rotations(rng, n_markers)
        mk_t = rng.normal(0.0, 0.3, (n_markers, 3))
    cam, time, marker = _visibility(rng, n_cams, n_times, n_markers, cams_per_t, marks_per_cam)
    E = cam.shape[0]
    # T_0^-1 . T_m  per marker
    R0T = mk_R[0].T
    rel_R = R0T[None] @ mk_R
    rel_t = (mk_t - mk_t[0]) @ R0T.T
    # T_t . (T_0^-1 T_m)
    wR = obj_R[time] @ rel_R[marker]
    wt = np.einsum("eij,ej->ei", obj_R[time], rel_t[marker]) + obj_t[time]
    # T_c^-1 . (...)
    RcT = np.transpose(cam_R[cam], (0, 2, 1))
    R = RcT @ wR
    t = np.einsum("eij,ej->ei", RcT, wt - cam_t[cam])
    R = R @ so3_exp(rng.normal(0.0, sigma_R, (E, 3))) if sigma_R > 0 else R
    t = t + rng.normal(0.0, sigma_t, (E, 3))
    reproj = rng.uniform(0.0, 0.01, E)
    if outlier_frac > 0:
        bad = rng.random(E) < outlier_frac
        nb = int(bad.sum())
        R[bad] = random_rotations(rng, nb)
        t[bad] = rng.normal(0.0, 5.0, (nb, 3))
        reproj[bad] = 1.0
    w = rng.uniform(0.5, 1.5, E)
    return SyntheticGraph(n_cams, n_times, n_markers, cam, time, marker, R, t, w, reproj,
                          mk_R, mk_t, cam_R, cam_t, obj_R, obj_t, "network")


def make_object_calibration(seed: int, n_times: int, n_markers: int = 24,
                            min_visible: int = 4, max_visible: int = 24,
                            sigma_R: float = 0.01, sigma_t: float = 0.005) -> SyntheticGraph:
    """cube_calib-shaped input of ``object_bipartite_se3sync`` (bipgo.py:493-545):
    one static camera, the object moves; key ``(str(t), f"{t}_{m}")``, pose =
    T_obj,t . T_m with the object about 4 m in front of the camera."""
    rng = np.random.default_rng(seed)
    mk_R, mk_t = _cube_markers() if n_markers == 24 else (
        random_rotations(rng, n_markers), rng.normal(0.0, 0.3, (n_markers, 3)))
    obj_R = random_rotations(rng, n_times)
    obj_t = rng.normal(0.0, 0.5, (n_times, 3)) + np.array([0.0, 0.0, 4.0])
    nvis = rng.integers(min_visible, max_visible + 1, n_times)
    mu = rng.random((n_times, n_markers))
    order = np.argsort(mu, axis=1)
    tt, mm = [], []
    for t in range(n_times):
        ms = np.sort(order[t, :nvis[t]])
        tt.append(np.full(ms.shape[0], t))
        mm.append(ms)
    time = np.concatenate(tt).astype(np.int64)
    marker = np.concatenate(mm).astype(np.int64)
    E = time.shape[0]
    R = obj_R[time] @ mk_R[marker]
    t = np.einsum("eij,ej->ei", obj_R[time], mk_t[marker]) + obj_t[time]
    R = R @ so3_exp(rng.normal(0.0, sigma_R, (E, 3)))
    t = t + rng.normal(0.0, sigma_t, (E, 3))
    reproj = rng.uniform(0.0, 0.01, E)
    w = rng.uniform(0.5, 1.5, E)
    return SyntheticGraph(1, n_times, n_markers, time.copy(), time, marker, R, t, w, reproj,
                          mk_R, mk_t, np.eye(3)[None], np.zeros((1, 3)), obj_R, obj_t, "object")


def to_edge_dict(g: SyntheticGraph, se3_cls) -> Tuple[Dict, Optional[Dict]]:
    """Build ``src_edges`` (and ``constraints`` for network graphs) with pose
    objects of class ``se3_cls`` (the reference's ``SE3`` or ours)."""
    edges = {}
    side = np.sqrt(g.w)
    cam, time, marker = g.cam.tolist(), g.time.tolist(), g.marker.tolist()
    for e in range(g.n_edges):
        c, t, m = cam[e], time[e], marker[e]
        key = (str(c), "%d_%d" % (t, m))
        edges[key] = {
            "pose": se3_cls(R=g.R[e].copy(), t=g.t[e].copy()),
            "corners": np.array([[0.0, 0.0], [side[e], 0.0], [side[e], side[e]], [0.0, side[e]]]),
            "reprojected_err": float(g.reproj[e]),
            "im_filename": "%d/%d.jpg" % (t, c),
            "w": float(g.w[e]),
        }
    if g.kind == "object":
        return edges, None
    constraints = {str(m): se3_cls(R=g.marker_R[m].copy(), t=g.marker_t[m].copy())
                   for m in range(g.n_markers)}
    return edges, constraints


def corner_area(corners: np.ndarray) -> float:
    """Shoelace area of the detected marker quadrilateral -- the quantity the notebook's noise
    models are built from (main.ipynb cells 3 and 7).  Plain Python arithmetic: it is called once
    or twice per detection by the synthetic noise models."""
    (x0, y0), (x1, y1), (x2, y2), (x3, y3) = corners.tolist()
    return 0.5 * abs((x0 * y1 - x1 * y0) + (x1 * y2 - x2 * y1) + (x2 * y3 - x3 * y2) + (x3 * y0 - x0 * y3))


def default_callables() -> Tuple[Callable, Callable, Callable]:
    """(noise_model_r, noise_model_t, edge_filter) of the synthetic recipe.  The
    weights are read from ``'corners'`` because ``object_bipartite_se3sync`` only
    forwards 'pose', 'corners', 'reprojected_err', 'im_filename' (bipgo.py:528-531);
    the synthetic quadrilateral is a square of area ``w``."""
    return ((lambda e: corner_area(e["corners"])),
            (lambda e: 2.0 * corner_area(e["corners"])),
            (lambda e: e["reprojected_err"] < 0.5))


# BASELINE.json configs (SURVEY.md 8d "Concrete configs").  cfg4 is generated on
# device by vican_b200.synthetic_device (dict API impossible at 50 M edges).
CONFIGS = {
    "cfg1": dict(fn="network", seed=1, n_cams=20, n_times=5000, n_markers=6, cams_per_t=7,
                 marks_per_cam=3, sigma_R=0.02, sigma_t=0.01, lsqr_solver="direct", maxiter=10),
    "cfg2": dict(fn="object", seed=0, n_times=2000, n_markers=24, sigma_R=0.01, sigma_t=0.005,
                 lsqr_solver="conjugate_gradient", maxiter=4),
    "cfg3": dict(fn="network", seed=11, n_cams=200, n_times=10000, n_markers=24, cams_per_t=20,
                 marks_per_cam=10, sigma_R=0.02, sigma_t=0.01, cube=True,
                 lsqr_solver="conjugate_gradient", maxiter=10),
    "cfg5": dict(fn="network", seed=11, n_cams=200, n_times=10000, n_markers=24, cams_per_t=20,
                 marks_per_cam=10, sigma_R=0.02, sigma_t=0.01, cube=True, outlier_frac=0.2,
                 lsqr_solver="conjugate_gradient", maxiter=500),
}


def make_config(name: str, scale: float = 1.0) -> Tuple[SyntheticGraph, dict]:
    """Instantiate a named config (optionally with n_times scaled down)."""
    cfg = dict(CONFIGS[name])
    fn = cfg.pop("fn")
    solver = cfg.pop("lsqr_solver")
    maxiter = cfg.pop("maxiter")
    cfg["n_times"] = max(8, int(round(cfg["n_times"] * scale)))
    g = make_camera_network(**cfg) if fn == "network" else make_object_calibration(**cfg)
    return g, dict(lsqr_solver=solver, maxiter=maxiter)
