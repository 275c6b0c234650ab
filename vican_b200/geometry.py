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

__all__ = ["SE3", "project_SO3", "geodesic_rad", "rel_translation_err"]


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
