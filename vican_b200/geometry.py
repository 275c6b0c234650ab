"""Host-side mirror of the reference's rigid-transform container.

Follows the *behaviour* of ``vican/geometry.py:194-261`` (class ``SE3``) and
``vican/geometry.py:175-191`` (``project_SO3``) so that objects produced by this
package can be passed to code written against the reference and vice versa:

* ``SE3(pose=M)`` rounds the 4x4 to float32 and ``R()``/``t()`` become float32
  views into it (reference ``geometry.py:208-211``).
* ``SE3(R=..., t=...)`` keeps the caller's arrays (and dtype) for ``R()``/``t()``
  and only the cached 4x4 is float32 (reference ``geometry.py:212-218``).
* ``inv()`` and ``@`` go through the float32 4x4 (reference ``geometry.py:235-243``,
  ``:260-261``) -- this rounding is part of the parity contract.

This is a container only; batch numerics live in the CUDA extension
(``vican_b200.ops.se3_compose_batch`` / ``se3_invert_batch`` / ``polar_so3_batch``).
"""
from __future__ import annotations

import numpy as np

__all__ = ["SE3", "project_SO3", "geodesic_rad", "rel_translation_err", "angle", "distance_SO3", "optimize_gauge_SO3",
           "optimize_gauge_SE3", "evaluate_against"]


class SE3:
    """3D rigid transformation (drop-in for the reference container)."""

    __slots__ = ("_pose", "_R", "_t")

    def __init__(self, **kwargs):
        if "pose" in kwargs:
            m = np.asarray(kwargs["pose"]).astype(np.float32)
            self._pose = m
            self._R = m[:3, :3]
            self._t = m[:3, -1]
        else:
            self._R = kwargs["R"]
            self._t = np.asarray(kwargs["t"]).flatten()
            m = np.zeros((4, 4), dtype=np.float32)
            m[:3, :3] += self._R
            m[:3, -1] += self._t
            m[-1, -1] += 1.0
            self._pose = m

    @classmethod
    def _from_pose32(cls, m: np.ndarray) -> "SE3":
        """``SE3(pose=m)`` for a float32 4x4 array the caller hands over (no copy: rows of a freshly built batch)."""
        self = cls.__new__(cls)
        self._pose = m
        self._R = m[:3, :3]
        self._t = m[:3, -1]
        return self

    def __getstate__(self):
        return {"_pose": self._pose, "_R": self._R, "_t": self._t}

    def __setstate__(self, state):
        # also accepts the attribute dictionary of a pickled reference SE3 (cam_marker_edges.pt)
        if isinstance(state, tuple):
            state = {**(state[0] or {}), **(state[1] or {})}
        for k in ("_pose", "_R", "_t"):
            object.__setattr__(self, k, state[k])

    def R(self) -> np.ndarray:
        return self._R

    def t(self) -> np.ndarray:
        return self._t

    def inv(self) -> "SE3":
        # float32 result buffer; products are evaluated in the dtype of _R/_t
        # and rounded on accumulation, as in the reference (geometry.py:239-243).
        out = np.zeros_like(self._pose)
        out[-1, -1] += 1
        out[:3, :3] += self._R.T
        out[:3, -1] += -self._R.T @ self._t
        return SE3(pose=out)

    def apply(self, x: np.ndarray) -> np.ndarray:
        assert x.ndim == 2
        assert x.shape[0] == 3
        return self._R @ x + self._t.reshape([-1, 1])

    def __matmul__(self, other: "SE3") -> "SE3":
        return SE3(pose=self._pose @ other._pose)

    def __repr__(self) -> str:
        return str(np.round(self._pose, 4))


def project_SO3(x: np.ndarray) -> np.ndarray:
    """Nearest rotation (host, numpy). Device version: ``ops.polar_so3_batch``."""
    u, _, vh = np.linalg.svd(x)
    d = np.linalg.det(u @ vh)
    return u @ np.diag([1.0, 1.0, d]) @ vh


def geodesic_rad(Ra: np.ndarray, Rb: np.ndarray) -> np.ndarray:
    """Geodesic angle (radians) between batches of rotations, accurate for tiny
    angles (uses the skew part, not arccos of the trace)."""
    Ra = np.asarray(Ra, dtype=np.float64).reshape(-1, 3, 3)
    Rb = np.asarray(Rb, dtype=np.float64).reshape(-1, 3, 3)
    D = np.einsum("nji,njk->nik", Ra, Rb)
    sk = np.stack([D[:, 2, 1] - D[:, 1, 2], D[:, 0, 2] - D[:, 2, 0], D[:, 1, 0] - D[:, 0, 1]], -1)
    s = 0.5 * np.linalg.norm(sk, axis=-1)
    c = 0.5 * (np.trace(D, axis1=1, axis2=2) - 1.0)
    return np.arctan2(s, c)


def rel_translation_err(ta: np.ndarray, tb: np.ndarray) -> np.ndarray:
    """Per-node relative translation error ||ta-tb|| / max(||tb||, 1e-12)."""
    ta = np.asarray(ta, dtype=np.float64).reshape(-1, 3)
    tb = np.asarray(tb, dtype=np.float64).reshape(-1, 3)
    return np.linalg.norm(ta - tb, axis=-1) / np.maximum(np.linalg.norm(tb, axis=-1), 1e-12)


# ---- evaluation helpers with the reference's signatures (geometry.py:131-172, :264-324); the
# ---- arithmetic runs in the CUDA extension (vican_b200.ops), there is no host fallback
def angle(r: np.ndarray) -> float:
    """Rotation angle of a 3x3 rotation in degrees (reference ``angle``, geometry.py:131-151)."""
    from . import ops
    return float(ops.distance_so3_batch(np.asarray(r, dtype=np.float64).reshape(1, 3, 3)).cpu()[0])


def distance_SO3(r1: np.ndarray, r2: np.ndarray) -> float:
    """Angle between two rotations in degrees (reference ``distance_SO3``, geometry.py:154-172)."""
    from . import ops
    r1, r2 = np.asarray(r1), np.asarray(r2)
    assert r1.shape == (3, 3) and r2.shape == (3, 3)
    return float(ops.distance_so3_batch(r1.reshape(1, 3, 3), r2.reshape(1, 3, 3)).cpu()[0])


def optimize_gauge_SO3(poses_a, poses_b) -> np.ndarray:
    """Rotation aligning ``poses_a`` with ``poses_b @ gauge`` (reference geometry.py:264-291)."""
    from . import ops
    assert len(poses_a) == len(poses_b)
    gR, _ = ops.optimize_gauge_batch(np.array([np.asarray(a, np.float64) for a in poses_a]), None,
                                     np.array([np.asarray(b, np.float64) for b in poses_b]), None)
    return gR.cpu().numpy()


def optimize_gauge_SE3(poses_a, poses_b) -> "SE3":
    """Transform aligning ``poses_a`` with ``poses_b @ gauge`` (reference geometry.py:294-324)."""
    from . import ops
    assert len(poses_a) == len(poses_b)
    gR, gt = ops.optimize_gauge_batch(np.array([np.asarray(a.R(), np.float64) for a in poses_a]),
                                      np.array([np.asarray(a.t(), np.float64) for a in poses_a]),
                                      np.array([np.asarray(b.R(), np.float64) for b in poses_b]),
                                      np.array([np.asarray(b.t(), np.float64) for b in poses_b]))
    return SE3(R=gR.cpu().numpy(), t=gt.cpu().numpy().reshape(3, 1))       # gauge_t is 3x1 in the reference (:315)


def evaluate_against(gt: dict, est: dict):
    """main.ipynb cell 9 as one device batch: gauge alignment of the estimates to the ground
    truth over the common keys, then per-key rotation error (degrees) and translation error
    (input units).  ``gt`` / ``est`` map ids to SE3-likes (poses wrt the world).  Returns
    (keys, r_err_deg [n], t_err [n], G) with numpy arrays and G the gauge as ``SE3``."""
    from . import ops
    keys = [k for k in gt.keys() if k in est]
    Rg = np.array([np.asarray(gt[k].R(), np.float64) for k in keys])
    tg = np.array([np.asarray(gt[k].t(), np.float64).reshape(3) for k in keys])
    Re = np.array([np.asarray(est[k].R(), np.float64) for k in keys])
    te = np.array([np.asarray(est[k].t(), np.float64).reshape(3) for k in keys])
    # cell 9 aligns the INVERSES: G = optimize_gauge_SE3([gt.inv()], [est.inv()]); est' = G.inv() @ est
    Rgi, tgi = ops.se3_invert_batch(Rg, tg)
    Rei, tei = ops.se3_invert_batch(Re, te)
    GR, Gt = ops.optimize_gauge_batch(Rgi, tgi, Rei, tei)
    GiR, Git = ops.se3_invert_batch(GR.reshape(1, 3, 3), Gt.reshape(1, 3))
    Ra, ta = ops.se3_left_compose_batch(GiR[0], Git[0], Re, te)
    r_err = ops.distance_so3_batch(Rg, Ra).cpu().numpy()
    import torch
    t_err = torch.linalg.vector_norm(torch.as_tensor(tg, device=ta.device) - ta, dim=1).cpu().numpy()
    return keys, r_err, t_err, SE3(R=GR.cpu().numpy(), t=Gt.cpu().numpy())
