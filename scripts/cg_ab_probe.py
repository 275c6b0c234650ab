"""A/B of the two CG mat-vec paths (VICAN_B200_CG_SMEM=0|1): translation time on cfg4 and bitwise comparison of the
results (both evaluate the same row sums in the same order).  Run once per setting; the second run compares."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
from vican_b200 import solver, synthetic as syn
from vican_b200.synthetic_device import make_scaled_network
tag = os.environ.get("VICAN_B200_CG_SMEM", "1")
dev = torch.device("cuda", 0)
out = {}
# cfg4
n_c, n_t = 10000, 1000000
det = make_scaled_network(4, n_c, n_t, 50, 0, n_t, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9); q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
g = solver.DeviceGraph(det.cam, det.time, det.marker, det.R, det.k_r, det.k_t, I9, n_c, n_t)
rot = solver.solve_rotations(g, 10)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr = solver.solve_translations(g, rot, det.t, q0, "conjugate_gradient")
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
print("CG_SMEM=%s cfg4: translation %.2f ms, %d iterations" % (tag, ms, tr.iters))
out["cfg4_xc"], out["cfg4_xt"] = tr.x_c.cpu(), tr.x_t.cpu()
del g, det, rot
# object-calibration shape through the array API and a ragged network
for name, s in (("obj", syn.make_object_calibration(0, 2000, 24)), ("net", syn.make_camera_network(3, 37, 901, 5, 9, 2, outlier_frac=0.1))):
    if name == "obj":
        cam, tm = s.marker.astype(np.int32), s.time.astype(np.int32); nc, nt = 24, 2000
        R = np.transpose(s.R, (0, 2, 1)).copy(); t = -np.einsum("eij,ej->ei", R, s.t); mk = np.zeros(cam.shape[0], np.int32); C = np.eye(3).reshape(1, 9); mq = np.zeros((1, 3))
    else:
        cam, tm, mk, R, t, nc, nt = s.cam.astype(np.int32), s.time.astype(np.int32), s.marker.astype(np.int32), s.R, s.t, 37, 901
        C = (np.transpose(s.marker_R, (0, 2, 1)) @ s.marker_R[0]).reshape(-1, 9); mq = np.zeros((5, 3))
    gg = solver.DeviceGraph(cam, tm, mk, R.reshape(-1, 9), s.w, 2.0 * s.w, C, nc, nt)
    rr = solver.solve_rotations(gg, 4)
    unk = (np.arange(nc, dtype=np.int32) * 3 % 7919 + 0, None)
    tt = solver.solve_translations(gg, rr, t, mq, "conjugate_gradient")
    out[name + "_xc"], out[name + "_xt"], out[name + "_it"] = tt.x_c.cpu(), tt.x_t.cpu(), tt.iters
    print("CG_SMEM=%s %s: %d iterations" % (tag, name, tt.iters))
os.makedirs("gpurun_out", exist_ok=True)
torch.save(out, "gpurun_out/cg_ab_%s.pt" % tag)
other = "gpurun_out/cg_ab_%s.pt" % ("0" if tag != "0" else "1")
if os.path.exists(other):
    o = torch.load(other)
    for k in out:
        same = torch.equal(out[k], o[k]) if isinstance(out[k], torch.Tensor) else out[k] == o[k]
        print("  %s identical: %s" % (k, same))
