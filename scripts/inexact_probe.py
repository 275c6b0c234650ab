"""Inexact inner solves: loosen the eigen tolerance of the EARLY outer iterations (tol_early, the last
`early_margin` iterations stay at 1e-13) and measure steps, time and the deviation of the final poses from the
all-tight solve, on cfg4 and on an outlier-laden large_shop-shaped graph."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np, torch
from vican_b200 import solver, synthetic as syn
from vican_b200.synthetic_device import make_scaled_network
from util import geodesic_rad, rel_translation_err
dev = torch.device("cuda", 0)

def run_case(name, g, t, marker_q, maxiter, margins=(4,)):
    base = None
    for tol_early, margin in [(0.0, 4)] + [(te, m) for te in (1e-5, 1e-4, 1e-3, 1e-2) for m in margins]:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rot = solver.solve_rotations(g, maxiter, tol_early=tol_early, early_margin=margin)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
        tr = solver.solve_translations(g, rot, t, marker_q, "conjugate_gradient")
        rc, rt = rot.r_c.cpu().numpy(), rot.r_t.cpu().numpy()
        xc, xt = tr.x_c.cpu().numpy(), tr.x_t.cpu().numpy()
        if base is None:
            base = (rc, rt, xc, xt)
        d = (geodesic_rad(rc, base[0]).max(), geodesic_rad(rt, base[1]).max(), rel_translation_err(xc, base[2]).max(),
             rel_translation_err(xt, base[3]).max())
        print("%s tol_early %.0e margin %d: steps %s (%d passes) rot %.2f ms | dev vs tight: rot %.1e / %.1e  trans %.1e / %.1e  cg %d%s"
              % (name, tol_early, margin, list(rot.stats.inner_per_outer[:min(maxiter, 12)]), rot.stats.time_passes + rot.stats.cam_passes, ms, *d, tr.iters,
                 "  REPEATED TIGHT" if getattr(rot, "repeated_tight", False) else ""), flush=True)

n_c, n_t = 10000, 1000000
det = make_scaled_network(4, n_c, n_t, 50, 0, n_t, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9); q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
g = solver.DeviceGraph(det.cam, det.time, det.marker, det.R, det.k_r, det.k_t, I9, n_c, n_t)
solver.solve_rotations(g, 2)
run_case("cfg4", g, det.t, q0, 10, margins=(3, 4))
del g, det
# cfg3 / cfg5 shape with 10 % outliers left IN (slow outer convergence), 20 and 60 iterations
s = syn.make_camera_network(11, 200, 3000, 24, 20, 10, outlier_frac=0.1, cube=True)
C_m = np.transpose(s.marker_R, (0, 2, 1)) @ s.marker_R[0]
g2 = solver.DeviceGraph(s.cam.astype(np.int32), s.time.astype(np.int32), s.marker.astype(np.int32), s.R.reshape(-1, 9), s.w, 2.0 * s.w, C_m, 200, 3000)
run_case("outliers-in maxiter 20", g2, s.t, np.zeros((24, 3)), 20, margins=(3, 4))
run_case("outliers-in maxiter 8", g2, s.t, np.zeros((24, 3)), 8, margins=(3, 4))
