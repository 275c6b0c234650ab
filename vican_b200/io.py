"""Data formats either side of the solver.

* ``load_edges``: reads the ``cam_marker_edges.pt`` files written by the reference's notebook
  (``torch.save(estimate_pose_mp(...))``, main.ipynb cells 3 and 5).  They are pickles of
  ``{(camera_id, "<timestamp>_<marker_id>"): {'pose': vican.geometry.SE3, 'corners': ...,
  'reprojected_err': ..., 'im_filename': ...}}`` (cam.py:176-184); the ``vican`` package need
  not be installed: its ``SE3`` is mapped onto this package's container while unpickling.
* ``EdgeAccumulator``: incremental flatten of per-image detection dictionaries as
  ``estimate_pose_worker`` produces them (cam.py:101-184), so that the solve after the last
  image starts from arrays (``EdgeTable.from_arrays``) instead of re-walking a dictionary.
* ``DeviceStream``: the same stream appended to a device-resident graph chunk by chunk
  (``solver.StreamingGraph``): no re-sort and no re-upload of what is already on the device.
"""
from __future__ import annotations

import gc
import pickle
from typing import Callable, Dict, Iterable, Optional

from .geometry import SE3

__all__ = ["load_edges", "save_edges", "EdgeAccumulator", "DeviceStream"]


class _RemapUnpickler(pickle.Unpickler):
    """``vican.geometry.SE3`` -> ``vican_b200.geometry.SE3`` (same attributes: _pose, _R, _t)."""

    def find_class(self, module, name):
        if module in ("vican.geometry", "geometry") and name == "SE3":
            return SE3
        return super().find_class(module, name)


class _RemapPickle:
    """``pickle_module`` for ``torch.load``: the stock module with the remapping unpickler."""
    __name__ = "vican_b200_remap_pickle"
    Unpickler = _RemapUnpickler
    load = staticmethod(lambda f, **kw: _RemapUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    Pickler = pickle.Pickler
    PicklingError = pickle.PicklingError
    UnpicklingError = pickle.UnpicklingError
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    DEFAULT_PROTOCOL = pickle.DEFAULT_PROTOCOL


def load_edges(path: str) -> dict:
    """Edge dictionary from a ``cam_marker_edges.pt`` (reference notebook) or from a file written
    by ``save_edges``; poses come back as this package's ``SE3``."""
    import torch
    gc_was_on = gc.isenabled()
    gc.disable()            # millions of small objects: keep the cyclic collector out of the load
    try:
        return torch.load(path, map_location="cpu", pickle_module=_RemapPickle, weights_only=False)
    finally:
        if gc_was_on:
            gc.enable()


def save_edges(edges: dict, path: str) -> None:
    """Counterpart of the notebook's ``torch.save(edges, path)``."""
    import torch
    torch.save(edges, path)


class EdgeAccumulator:
    """Streaming front end of the solver: feed the per-image dictionaries of
    ``estimate_pose_worker`` (or any ``{(cam, "t_m"): {...}}`` chunk) as they arrive; the
    callables run once per detection at ``add`` time, and ``table(constraints)`` hands the
    accumulated arrays to the device ingestion without touching the detections again."""

    def __init__(self, noise_model_r: Callable, noise_model_t: Callable, edge_filter: Callable,
                 marker_ids: Optional[Iterable[str]] = None):
        self.noise_model_r, self.noise_model_t, self.edge_filter = noise_model_r, noise_model_t, edge_filter
        self.marker_ids = None if marker_ids is None else set(marker_ids)     # cam.py:263 id whitelist
        self._cams, self._tm, self._R, self._t, self._kr, self._kt = [], [], [], [], [], []
        self._seen: Dict[tuple, int] = {}
        self._dead = set()              # positions whose key was later re-detected and then filtered out
        self.n_dropped = 0

    def __len__(self) -> int:
        return len(self._cams) - len(self._dead)

    def add(self, detections: Optional[dict]) -> int:
        """Returns the number of detections kept from this chunk.  A key seen before replaces the
        older detection (dictionary-merge semantics of cam.py:263) -- also when the newer detection
        fails ``edge_filter``: the merged dictionary would hold the newer one, which the solver then
        filters out, so the older record is retired."""
        if not detections:
            return 0
        kept = 0
        for key, v in detections.items():
            if self.marker_ids is not None and key[-1].split("_")[-1] not in self.marker_ids:
                continue
            pos = self._seen.get(key)
            if not self.edge_filter(v):
                self.n_dropped += 1
                if pos is not None:
                    self._dead.add(pos)
                continue
            pose = v["pose"]
            rec = (key[0], key[1], pose.R(), pose.t(), self.noise_model_r(v), self.noise_model_t(v))
            if pos is not None:
                self._dead.discard(pos)
            if pos is None:
                self._seen[key] = len(self._cams)
                for lst, x in zip((self._cams, self._tm, self._R, self._t, self._kr, self._kt), rec):
                    lst.append(x)
            else:
                for lst, x in zip((self._cams, self._tm, self._R, self._t, self._kr, self._kt), rec):
                    lst[pos] = x
            kept += 1
        return kept

    def table(self, constraints: dict):
        from .bipgo import EdgeTable
        tab = EdgeTable.__new__(EdgeTable)
        cols = (self._cams, self._tm, self._R, self._t, self._kr, self._kt)
        if self._dead:
            cols = tuple([x for i, x in enumerate(col) if i not in self._dead] for col in cols)
        tab._assemble(list(cols[0]), list(cols[1]), list(cols[2]), list(cols[3]), list(cols[4]), list(cols[5]),
                      constraints)
        return tab


class DeviceStream:
    """Streaming ingestion onto the device: per-image detection dictionaries, as
    ``estimate_pose_worker`` returns them (cam.py:101-184), are flattened on arrival and appended to a
    ``solver.StreamingGraph`` in chunks of COMPLETE timesteps -- the resident block-CSR is never re-sorted.

    The camera set is fixed up front (a deployed network knows its cameras; index order = the reference's
    lexicographic order, so index 0 is the gauge camera).  Timesteps must arrive in order: a timestep is
    appended once a later one has been seen (``flush(final=True)`` appends the rest), and detections that
    arrive for an already appended timestep are counted in ``n_late`` and dropped.  Time nodes are numbered
    in arrival order, not in the reference's lexicographic order, so results agree with the reference to
    rounding (<= 1e-9 rad; translations as far as the truncated CG is stable), not bit for bit."""

    def __init__(self, camera_ids, constraints: dict, noise_model_r: Callable, noise_model_t: Callable,
                 edge_filter: Callable, chunk_detections: int = 4096, device=None):
        import numpy as np
        from . import solver
        self._np, self._solver = np, solver
        self.cam_ids = np.unique(np.asarray([str(c) for c in camera_ids]))
        self._cpos = {str(c): i for i, c in enumerate(self.cam_ids)}
        self.constraints = constraints
        self.root = str(min(list(constraints.keys())))                      # bipgo.py:411
        self.marker_ids = sorted(str(m) for m in constraints.keys())
        self._mpos = {m: i for i, m in enumerate(self.marker_ids)}
        R0 = np.asarray(constraints[self.root].R(), dtype=np.float64)
        markerC = np.stack([np.asarray(constraints[m].R(), dtype=np.float64).T @ R0 for m in self.marker_ids])
        q = []
        for m in self.marker_ids:                                           # bipgo.py:451-452
            r_0m = constraints[self.root].R().T @ constraints[m].R()
            t_m0 = (constraints[m].inv() @ constraints[self.root]).t()
            q.append(np.asarray(r_0m, dtype=np.float64) @ np.asarray(t_m0, dtype=np.float64))
        self.marker_q = np.stack(q)
        self.graph = solver.StreamingGraph(len(self.cam_ids), markerC, device=device)
        self.noise_model_r, self.noise_model_t, self.edge_filter = noise_model_r, noise_model_t, edge_filter
        self.chunk_detections = int(chunk_detections)
        self.time_ids = []                  # appended timesteps, in device index order
        self._done = set()
        self._pending: Dict[str, list] = {}  # timestep -> detections (arrival order)
        self._n_pending = 0
        self.n_late = self.n_dropped = 0

    def add(self, detections: Optional[dict]) -> None:
        if not detections:
            return
        for key, v in detections.items():
            ts, mk = key[1].split("_")
            if ts in self._done:
                self.n_late += 1
                continue
            if not self.edge_filter(v):
                self.n_dropped += 1
                continue
            pose = v["pose"]
            self._pending.setdefault(ts, []).append((self._cpos[str(key[0])], self._mpos[mk], pose.R(), pose.t(),
                                                     self.noise_model_r(v), self.noise_model_t(v)))
            self._n_pending += 1
        if self._n_pending >= self.chunk_detections and len(self._pending) > 1:
            self.flush(final=False)

    def flush(self, final: bool = True) -> None:
        """Append the buffered timesteps (all but the newest unless ``final``) as one chunk."""
        np = self._np
        keys = list(self._pending.keys())
        if not final:
            keys = keys[:-1]
        if not keys:
            return
        cam, time, marker, R, t, kr, kt = [], [], [], [], [], [], []
        for j, ts in enumerate(keys):
            for (c, m, Rm, tv, a, b) in self._pending.pop(ts):
                cam.append(c); time.append(j); marker.append(m); R.append(Rm); t.append(tv); kr.append(a); kt.append(b)
            self._done.add(ts)
            self.time_ids.append(ts)
        self._n_pending -= len(cam)
        self.graph.append(np.asarray(cam, dtype=np.int32), np.asarray(time, dtype=np.int32), np.asarray(marker, dtype=np.int32),
                          np.asarray(R, dtype=np.float64).reshape(-1, 9), np.asarray(t, dtype=np.float64).reshape(-1, 3),
                          np.asarray(kr, dtype=np.float64), np.asarray(kt, dtype=np.float64), len(keys))

    def solve(self, maxiter: int, lsqr_solver: str = "conjugate_gradient", dtype=None) -> dict:
        """``bipartite_se3sync`` on everything streamed so far: {camera id / f"{t}_0": SE3}."""
        np, solver = self._np, self._solver
        self.flush(final=True)
        g = self.graph.graph()
        rot = solver.solve_rotations(g, maxiter)
        tr = solver.solve_translations(g, rot, self.graph.t, self.marker_q, lsqr_solver)
        Rc, Rt = rot.world_rotations()
        Rc, Rt = Rc.cpu().numpy(), Rt.cpu().numpy()
        if dtype is not None:
            Rc, Rt = Rc.astype(dtype), Rt.astype(dtype)
        xc, xt = tr.x_c.cpu().numpy(), tr.x_t.cpu().numpy()
        out = {str(c): SE3(R=Rc[i], t=xc[i]) for i, c in enumerate(self.cam_ids)}
        out.update({ts + "_0": SE3(R=Rt[j], t=xt[j]) for j, ts in enumerate(self.time_ids)})
        return out
