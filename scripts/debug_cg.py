"""Debug aid: device CG vs scipy cg on the object-calibration shape, array API (both unknown orders)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import scipy.sparse.linalg as spl
from oracle import vican_oracle as orc
from vican_b200 import synthetic as syn, solver
from util import rel_translation_err

for n_t in (1200, 2000, 500):
    g = syn.make_object_calibration(0, n_t, 24)
    cam, time = g.marker.astype(np.int32), g.time.astype(np.int32)
    zeros = np.zeros(cam.shape[0], dtype=np.int32)
    Rinv = np.transpose(g.R, (0, 2, 1)).copy()
    tinv = -np.einsum("eij,ej->ei", Rinv, g.t)
    k_r, k_t = g.w, 2.0 * g.w
    I9 = np.eye(3).reshape(1, 9)
    n_c = 24
    dg = solver.DeviceGraph(cam, time, zeros, Rinv, k_r, k_t, I9, n_c, n_t)
    rot = solver.solve_rotations(dg, 4)
    a = solver.solve_translations(dg, rot, tinv, np.zeros((1, 3)), "conjugate_gradient")
    Rw_c, Rw_t = rot.world_rotations()
    J, tt = orc.translation_system(cam.astype(np.int64), time.astype(np.int64), zeros.astype(np.int64), tinv, k_t,
                                   np.eye(3)[None], np.zeros((1, 3)), 0, Rw_c.cpu().numpy(), Rw_t.cpu().numpy(),
                                   n_c, n_t, np.arange(n_c), n_c + np.arange(n_t))
    cnt = [0]
    x, code = spl.cg(J.T @ J, J.T @ tt, callback=lambda xk: cnt.__setitem__(0, cnt[0] + 1))
    x = x.reshape(-1, 3)
    ec = rel_translation_err(a.x_c.cpu().numpy(), x[:n_c]).max()
    et = rel_translation_err(a.x_t.cpu().numpy(), x[n_c:]).max()
    print("n_t %d: device iters %d scipy iters %d  err cam %.3e time %.3e" % (n_t, a.iters, cnt[0], ec, et), flush=True)
    # rhs check
    b = (J.T @ tt).reshape(-1, 3)
