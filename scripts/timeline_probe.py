"""Kernel timeline of one cfg4 solve (torch profiler / CUPTI): busy time, idle gaps, per-kernel totals."""
import os, sys, collections
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch.profiler import profile, ProfilerActivity
from vican_b200 import solver
from vican_b200.synthetic_device import make_scaled_network
dev = torch.device("cuda", 0)
det = make_scaled_network(4, 10000, 1000000, 50, 0, 1000000, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9); q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
def run():
    return solver.solve_arrays(det.cam, det.time, det.marker, det.R, det.t, det.k_r, det.k_t, I9, q0, 10000, det.n_t, 10, "conjugate_gradient")
for _ in range(2): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = run(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy = 0.0; last_end = t0; gaps = []
agg = collections.OrderedDict()
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if s > last_end:
        gaps.append((s - last_end, e.name[:50]))
    busy += max(0.0, en - max(s, last_end)); last_end = max(last_end, en)
    a = agg.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += (en - s)
print("span %.2f ms, busy %.2f ms, idle %.2f ms, kernels %d, phases %s" % ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(evs), r.phase_ms))
gaps.sort(reverse=True)
print("largest gaps (us, next kernel):", [(round(g, 1), n) for g, n in gaps[:12]])
print("gap histogram: >100us %d, 20-100us %d, 5-20us %d, <5us %d; sum of gaps <20us: %.2f ms" % (
    sum(g > 100 for g, _ in gaps), sum(20 < g <= 100 for g, _ in gaps), sum(5 < g <= 20 for g, _ in gaps), sum(g <= 5 for g, _ in gaps),
    sum(g for g, _ in gaps if g <= 20) / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("%-62s n=%4d %8.3f ms avg %.4f" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0]))
