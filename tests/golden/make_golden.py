"""Generate golden vectors by running the REAL reference (``/root/reference/vican``)
in the build container.  Run once, here:  ``python tests/golden/make_golden.py``.

The GPU box has no /root/reference, so inputs (arrays) and the reference's outputs
are committed as small ``.npz`` files next to this script.  The reference's
non-symmetric ``eigs`` occasionally returns a complex pair and crashes with
``LinAlgError``; such runs are simply retried -- successful
runs are reproducible to 1e-15.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference")
os.environ["TQDM_DISABLE"] = "1"

from vican.bipgo import bipartite_se3sync, object_bipartite_se3sync  # noqa: E402  (reference)
from vican.geometry import SE3 as RefSE3                              # noqa: E402  (reference)

from vican_b200 import synthetic as syn                               # noqa: E402

CASES = {
    # name: (graph kwargs, solver kwargs, filter on?)
    "net_tiny_cg_it1": (dict(fn="network", seed=3, n_cams=5, n_times=40, n_markers=3, cams_per_t=3, marks_per_cam=2),
                        dict(maxiter=1, lsqr_solver="conjugate_gradient"), True),
    "net_small_cg_it3": (dict(fn="network", seed=5, n_cams=12, n_times=80, n_markers=4, cams_per_t=4, marks_per_cam=2),
                         dict(maxiter=3, lsqr_solver="conjugate_gradient"), True),
    "net_small_direct_it3": (dict(fn="network", seed=5, n_cams=12, n_times=80, n_markers=4, cams_per_t=4, marks_per_cam=2),
                             dict(maxiter=3, lsqr_solver="direct"), True),
    "net_small_cg_it2": (dict(fn="network", seed=6, n_cams=12, n_times=80, n_markers=4, cams_per_t=4, marks_per_cam=2),
                         dict(maxiter=2, lsqr_solver="conjugate_gradient"), True),
    "net_small_cg_it10": (dict(fn="network", seed=7, n_cams=12, n_times=80, n_markers=4, cams_per_t=4, marks_per_cam=2),
                          dict(maxiter=10, lsqr_solver="conjugate_gradient"), True),
    "net_outl_filtered_it5": (dict(fn="network", seed=8, n_cams=15, n_times=120, n_markers=6, cams_per_t=5,
                                   marks_per_cam=3, outlier_frac=0.2),
                              dict(maxiter=5, lsqr_solver="conjugate_gradient"), True),
    "net_outl_unfiltered_it5": (dict(fn="network", seed=9, n_cams=15, n_times=120, n_markers=6, cams_per_t=5,
                                     marks_per_cam=3, outlier_frac=0.1),
                                dict(maxiter=5, lsqr_solver="direct"), False),
    "net_medium_cg_it4": (dict(fn="network", seed=1, n_cams=20, n_times=300, n_markers=6, cams_per_t=7, marks_per_cam=3),
                          dict(maxiter=4, lsqr_solver="conjugate_gradient"), True),
    "net_cube_cg_it4": (dict(fn="network", seed=11, n_cams=30, n_times=150, n_markers=24, cams_per_t=8,
                             marks_per_cam=5, cube=True),
                        dict(maxiter=4, lsqr_solver="conjugate_gradient"), True),
    "obj_small_cg_it4": (dict(fn="object", seed=0, n_times=120, n_markers=24, sigma_R=0.01, sigma_t=0.005),
                         dict(maxiter=4, lsqr_solver="conjugate_gradient"), True),
    "obj_small_direct_it2": (dict(fn="object", seed=2, n_times=60, n_markers=10, min_visible=3, max_visible=10),
                             dict(maxiter=2, lsqr_solver="direct"), True),
}


def build_graph(gkw):
    gkw = dict(gkw)
    fn = gkw.pop("fn")
    return syn.make_camera_network(**gkw) if fn == "network" else syn.make_object_calibration(**gkw)


def run_reference(g, skw, filt_on):
    edges, constraints = syn.to_edge_dict(g, RefSE3)
    nr, nt, ef = syn.default_callables()
    if not filt_on:
        ef = lambda e: True  # noqa: E731
    last = None
    for attempt in range(8):
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                if g.kind == "object":
                    out = object_bipartite_se3sync(edges, noise_model_r=nr, noise_model_t=nt, edge_filter=ef,
                                                   dtype=np.float64, **skw)
                else:
                    out = bipartite_se3sync(edges, constraints=constraints, noise_model_r=nr, noise_model_t=nt,
                                            edge_filter=ef, dtype=np.float64, **skw)
            return out
        except np.linalg.LinAlgError as exc:  # reference's latent eigs crash -> retry
            last = exc
    raise last


def main():
    for name, (gkw, skw, filt_on) in CASES.items():
        g = build_graph(gkw)
        out = run_reference(g, skw, filt_on)
        keys = sorted(out.keys())
        R = np.stack([np.asarray(out[k].R(), dtype=np.float64) for k in keys])
        t = np.stack([np.asarray(out[k].t(), dtype=np.float64) for k in keys])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            path, kind=g.kind, n_cams=g.n_cams, n_times=g.n_times, n_markers=g.n_markers,
            cam=g.cam.astype(np.int32), time=g.time.astype(np.int32), marker=g.marker.astype(np.int32),
            R=g.R, t=g.t, w=g.w, reproj=g.reproj, marker_R=g.marker_R, marker_t=g.marker_t,
            maxiter=skw["maxiter"], lsqr_solver=skw["lsqr_solver"], filter_on=filt_on,
            out_keys=np.array(keys), out_R=R, out_t=t)
        print("%-28s E_raw=%6d  nodes=%5d  -> %s (%.0f KB)" % (name, g.n_edges, len(keys), os.path.basename(path),
                                                              os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
