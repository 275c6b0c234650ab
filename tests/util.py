"""Shared helpers for the parity tests (golden loading, metrics)."""
import glob
import os

import numpy as np

from vican_b200 import synthetic as syn
from vican_b200.geometry import SE3


# The checker's own metrics (kept out of the product package on purpose).
def geodesic_rad(Ra, Rb):
    """Geodesic distance on SO(3) between stacked rotations, in radians, evaluated through the norm of
    the skew part and the trace (atan2: accurate for tiny angles, where arccos of the trace loses half
    the digits)."""
    Ra = np.asarray(Ra, dtype=np.float64).reshape(-1, 3, 3)
    Rb = np.asarray(Rb, dtype=np.float64).reshape(-1, 3, 3)
    D = np.transpose(Ra, (0, 2, 1)) @ Rb
    s = 0.5 * np.sqrt((D[:, 2, 1] - D[:, 1, 2]) ** 2 + (D[:, 0, 2] - D[:, 2, 0]) ** 2 + (D[:, 1, 0] - D[:, 0, 1]) ** 2)
    c = 0.5 * (np.trace(D, axis1=1, axis2=2) - 1.0)
    return np.arctan2(s, c)


def rel_translation_err(ta, tb):
    """Per-node relative translation error |ta - tb| / max(|tb|, tiny)."""
    ta = np.asarray(ta, dtype=np.float64).reshape(-1, 3)
    tb = np.asarray(tb, dtype=np.float64).reshape(-1, 3)
    return np.linalg.norm(ta - tb, axis=1) / np.maximum(np.linalg.norm(tb, axis=1), 1e-300)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity tolerances stated by BASELINE.json north_star (fp64).
ROT_TOL_RAD = 1e-6
TRANS_REL_TOL = 1e-6


def golden_names():
    """Solver goldens (network / object cases); ``eval_*.npz`` belong to the evaluation helpers."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith(("net_", "obj_"))]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    kind = str(z["kind"])
    nm = int(z["n_markers"])
    g = syn.SyntheticGraph(
        int(z["n_cams"]), int(z["n_times"]), nm, z["cam"].astype(np.int64), z["time"].astype(np.int64),
        z["marker"].astype(np.int64), z["R"], z["t"], z["w"], z["reproj"], z["marker_R"], z["marker_t"],
        None, None, None, None, kind)
    params = dict(maxiter=int(z["maxiter"]), lsqr_solver=str(z["lsqr_solver"]))
    ref = {str(k): (z["out_R"][i], z["out_t"][i]) for i, k in enumerate(z["out_keys"])}
    return g, params, bool(z["filter_on"]), ref


def callables(filter_on=True):
    nr, nt, ef = syn.default_callables()
    if not filter_on:
        ef = lambda e: True  # noqa: E731
    return nr, nt, ef


def as_pairs(out):
    """{key: SE3-like or (R,t)} -> {str key: (R fp64, t fp64)}"""
    res = {}
    for k, v in out.items():
        if isinstance(v, tuple):
            res[str(k)] = (np.asarray(v[0], np.float64), np.asarray(v[1], np.float64))
        else:
            res[str(k)] = (np.asarray(v.R(), np.float64), np.asarray(v.t(), np.float64))
    return res


def compare(out, ref):
    """max geodesic error (rad) and max per-node relative translation error."""
    out, ref = as_pairs(out), as_pairs(ref)
    assert set(out.keys()) == set(ref.keys()), (sorted(set(out) ^ set(ref))[:10])
    keys = sorted(ref.keys())
    Ra = np.stack([out[k][0] for k in keys])
    Rb = np.stack([ref[k][0] for k in keys])
    ta = np.stack([out[k][1] for k in keys])
    tb = np.stack([ref[k][1] for k in keys])
    return float(geodesic_rad(Ra, Rb).max()), float(rel_translation_err(ta, tb).max())
