import os, sys, time, dataclasses
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from vican_b200 import solver
from vican_b200.synthetic_device import make_scaled_network
dev = torch.device("cuda", 0)
det = make_scaled_network(4, 10000, 1000000, 50, 0, 1000000, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9); q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
host = dataclasses.replace(det, **{f: getattr(det, f).cpu().pin_memory() for f in ("cam", "time", "marker", "R", "t", "k_r", "k_t")})
def run(src, to_host):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = solver.solve_arrays(src.cam, src.time, src.marker, src.R, src.t, src.k_r, src.k_t, I9, q0, 10000, src.n_t, 10,
                            "conjugate_gradient", to_host=to_host, reuse_host_buffers=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r.phase_ms
for i in range(3): print("device", run(det, False))
for i in range(3): print("host  ", run(host, True))
# raw H2D rate
torch.cuda.synchronize(); t0 = time.perf_counter(); x = host.R.to(dev, non_blocking=True); torch.cuda.synchronize()
dt = time.perf_counter() - t0; print("H2D R: %.1f ms  %.1f GB/s" % (dt * 1e3, host.R.numel() * 8 / dt / 1e9))
