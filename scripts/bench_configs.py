#!/usr/bin/env python
"""BASELINE.json configs 0-2 and 4 through the dict API (full size), with the CPU oracle next to it.

    python scripts/bench_configs.py [cfg1 cfg2 cfg3 cfg5] [--no-oracle]

Prints one JSON line per config: wall time of the drop-in call (dict flatten on the host + device
solve), the device part alone, the oracle's wall time on the same dict, and the parity errors."""
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    names = args or ["cfg1", "cfg2", "cfg3"]
    with_oracle = "--no-oracle" not in sys.argv
    import torch
    from oracle import vican_oracle as orc
    from vican_b200 import bipgo, synthetic as syn
    from vican_b200.geometry import SE3, geodesic_rad, rel_translation_err
    nr, nt, ef = syn.default_callables()
    # warm the device / library once
    g0 = syn.make_camera_network(0, 6, 20, 3, 3, 2)
    e0, c0 = syn.to_edge_dict(g0, SE3)
    bipgo.bipartite_se3sync(e0, c0, nr, nt, ef, 2, "conjugate_gradient")
    for name in names:
        g, p = syn.make_config(name)
        if name == "cfg5":
            p["maxiter"] = int(os.environ.get("CFG5_MAXITER", "500"))
        t0 = time.perf_counter()
        edges, cons = syn.to_edge_dict(g, SE3)
        t_dict = time.perf_counter() - t0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if g.kind == "object":
            out = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **p)
        else:
            out = bipgo.bipartite_se3sync(edges, cons, nr, nt, ef, dtype=np.float64, **p)
        t_ours = time.perf_counter() - t0
        info = dict(bipgo.last_info)
        rec = dict(config=name, kind=g.kind, n_raw=g.n_edges, n_c=info["n_c"], n_t=info["n_t"], n_edges=info["n_edges"],
                   maxiter=p["maxiter"], lsqr_solver=p["lsqr_solver"], build_dict_s=round(t_dict, 3),
                   ours_wall_s=round(t_ours, 4), ours_device_s=round(info["device_seconds"], 4),
                   inner_per_outer=info["inner_per_outer"][:12], trans_iters=info["trans_iters"])
        if with_oracle and p["maxiter"] <= 20:
            t0 = time.perf_counter()
            if g.kind == "object":
                ref = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, se3_cls=SE3, **p)
            else:
                ref = orc.bipartite_se3sync_oracle(edges, cons, nr, nt, ef, **p)
            rec["oracle_wall_s"] = round(time.perf_counter() - t0, 3)
            keys = sorted(ref.keys())
            Ra = np.stack([np.asarray(out[k].R(), np.float64) for k in keys]); Rb = np.stack([ref[k][0] for k in keys])
            ta = np.stack([out[k].t() for k in keys]); tb = np.stack([ref[k][1] for k in keys])
            rec["rot_err_rad"] = float(geodesic_rad(Ra, Rb).max())
            rec["rel_t_err"] = float(rel_translation_err(ta, tb).max())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
