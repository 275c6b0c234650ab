"""Golden OUTPUTS of the REAL reference (``/root/reference/vican``) on the BASELINE.json configurations at FULL
size: cfg1 (both solvers), cfg2, cfg3 -- and on cfg5 (maxiter 500) at 10 % of its time nodes.  Run once, in the build container:

    python tests/golden/make_golden_fullsize.py            # ~6 min (cfg3: 2 M detections through the reference)

The inputs are NOT stored (cfg3: 2 M detections): the synthetic recipe is seeded (``vican_b200.synthetic.make_config``,
numpy's PCG64 stream is stable across versions), so the tests regenerate them and check a SHA-256 digest of the
regenerated arrays against the one stored here before comparing any result.  Stored per case: node keys, R, t of the
reference's answer (float64) and the digest.  ``full_*.npz`` files are not picked up by ``util.golden_names()``.
"""
import contextlib
import hashlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference")
os.environ["TQDM_DISABLE"] = "1"

from vican.bipgo import bipartite_se3sync, object_bipartite_se3sync  # noqa: E402  (reference)
from vican.geometry import SE3 as RefSE3                              # noqa: E402  (reference)

from vican_b200 import synthetic as syn                               # noqa: E402

CASES = {   # name: (config, solver override, scale of the time nodes)
    "full_cfg1_direct": ("cfg1", "direct", 1.0),
    "full_cfg1_cg": ("cfg1", "conjugate_gradient", 1.0),
    "full_cfg2": ("cfg2", None, 1.0),
    "full_cfg3": ("cfg3", None, 1.0),
    "tenth_cfg5": ("cfg5", None, 0.1),      # maxiter = 500 with 20 % outliers behind edge_filter; the size the GPU test runs
}


def input_digest(g):
    """SHA-256 over the arrays the detection dictionary is built from (tests/util.py carries the same function)."""
    h = hashlib.sha256()
    for a in (g.cam.astype(np.int64), g.time.astype(np.int64), g.marker.astype(np.int64), g.R, g.t, g.w, g.reproj,
              g.marker_R, g.marker_t):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def run_reference(g, skw):
    edges, constraints = syn.to_edge_dict(g, RefSE3)
    nr, nt, ef = syn.default_callables()
    last = None
    for attempt in range(8):
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                if g.kind == "object":
                    return object_bipartite_se3sync(edges, noise_model_r=nr, noise_model_t=nt, edge_filter=ef,
                                                    dtype=np.float64, **skw)
                return bipartite_se3sync(edges, constraints=constraints, noise_model_r=nr, noise_model_t=nt,
                                         edge_filter=ef, dtype=np.float64, **skw)
        except np.linalg.LinAlgError as exc:  # the reference's latent eigs crash -> retry
            last = exc
    raise last


def main(only=None):
    for name, (cfg, solver, scale) in CASES.items():
        if only and name not in only:
            continue
        t0 = time.perf_counter()
        g, skw = syn.make_config(cfg, scale)
        if solver is not None:
            skw["lsqr_solver"] = solver
        out = run_reference(g, skw)
        keys = sorted(out.keys())
        R = np.stack([np.asarray(out[k].R(), dtype=np.float64) for k in keys])
        t = np.stack([np.asarray(out[k].t(), dtype=np.float64) for k in keys])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, config=cfg, maxiter=skw["maxiter"], lsqr_solver=skw["lsqr_solver"], kind=g.kind,
                            n_detections=g.n_edges, input_sha256=input_digest(g), out_keys=np.array(keys), out_R=R, out_t=t)
        print("%-18s %s E_raw=%7d nodes=%5d -> %s (%.0f KB), %.0f s" % (name, cfg, g.n_edges, len(keys), os.path.basename(path),
                                                                     os.path.getsize(path) / 1024, time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
