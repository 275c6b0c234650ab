"""Times the two edge passes alone on the cfg4 graph (CUDA events, 20 launches each, inputs >> L2)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from vican_b200 import _cabi, solver
from vican_b200.synthetic_device import make_scaled_network
lib = _cabi.lib()
dev = torch.device("cuda", 0)
n_c, n_t, d = 10000, 1000000, 50
det = make_scaled_network(4, n_c, n_t, d, 0, n_t, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
g = solver.DeviceGraph(det.cam, det.time, det.marker, det.R, det.k_r, det.k_t, I9, n_c, n_t)
gs = int(lib.vb_gather_stride())
X = torch.randn((n_c, gs), dtype=torch.float64, device=dev)
lamT = torch.randn((n_t, 9), dtype=torch.float64, device=dev)
Wt = torch.zeros((n_t, gs), dtype=torch.float64, device=dev)
Y = torch.zeros((n_c, 9), dtype=torch.float64, device=dev)
p, st = solver._ptr, solver._stream
for name, fn, kind in (("time", lambda: lib.vb_pass_time(C.byref(g.cgraph), 0, p(X), p(lamT), p(Wt), st()), "time"),
                       ("cam", lambda: lib.vb_pass_cam(C.byref(g.cgraph), p(Wt), p(Y), st()), "cam")):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    nb = g.pass_bytes(kind)
    print("%s pass: %.4f ms  %.0f GB/s  frac of 6461: %.3f" % (name, ms, nb / ms * 1e-6, nb / ms * 1e-6 / 6461.2))
print("checksum Wt %.15e  Y %.15e  var %s" % (float(Wt.double().abs().sum()), float(Y.abs().sum()), os.environ.get("VICAN_B200_PASS_VAR", "default")))
