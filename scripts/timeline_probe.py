"""Kernel timeline of one cfg4 solve (torch profiler / CUPTI) on rank 0: busy time, idle gaps, per-kernel totals.
Single GPU: python scripts/timeline_probe.py;  N GPUs: torchrun --nproc-per-node N scripts/timeline_probe.py"""
import os, sys, collections
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch.profiler import profile, ProfilerActivity
from vican_b200 import dist as vdist, solver
from vican_b200.synthetic_device import make_scaled_network
rank, world = vdist.init_process_group_from_env("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
comm = vdist.create_comm()
n_c, n_t = 10000, 1000000
lo, hi = vdist.shard_range(n_t, rank, world)
det = make_scaled_network(4, n_c, n_t, 50, lo, hi, device=dev)
I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9); q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
def run():
    return solver.solve_arrays(det.cam, det.time, det.marker, det.R, det.t, det.k_r, det.k_t, I9, q0, n_c, det.n_t, 10,
                               "conjugate_gradient", comm=comm)
for _ in range(3): run()
torch.cuda.synchronize()
if world > 1:
    import torch.distributed as tdist
    tdist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = run(); torch.cuda.synchronize()
if rank == 0:
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
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/timeline_%dgpu.csv" % world, "w") as f:
        f.write("start_us,dur_us,gap_before_us,name\n")
        prev = t0
        for e in evs:
            f.write("%.1f,%.1f,%.1f,%s\n" % (e.time_range.start - t0, e.time_range.end - e.time_range.start,
                                             e.time_range.start - prev, e.name[:70].replace(",", ";")))
            prev = max(prev, e.time_range.end)
    print("world %d: span %.2f ms, busy %.2f ms, idle %.2f ms, kernels %d, phases %s" % (world, (t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(evs), r.phase_ms))
    gaps.sort(reverse=True)
    print("largest gaps (us, next kernel):", [(round(g, 1), n) for g, n in gaps[:14]])
    print("gap histogram: >100us %d, 20-100us %d, 5-20us %d, <5us %d; sum of gaps <20us: %.2f ms, 20-100us: %.2f ms" % (
        sum(g > 100 for g, _ in gaps), sum(20 < g <= 100 for g, _ in gaps), sum(5 < g <= 20 for g, _ in gaps), sum(g <= 5 for g, _ in gaps),
        sum(g for g, _ in gaps if g <= 20) / 1e3, sum(g for g, _ in gaps if 20 < g <= 100) / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
        print("%-62s n=%4d %8.3f ms avg %.4f" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0]))
vdist.destroy_comm(comm)
if world > 1:
    tdist.destroy_process_group()
