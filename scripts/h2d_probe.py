"""Host-to-device bandwidth per rank when N ranks upload at once, with and without the pinned buffer placed on the
GPU's own NUMA node (the e2e arm of bench.py at N > 1 is bound by this).
    torchrun --nproc-per-node N scripts/h2d_probe.py"""
import os, sys, subprocess
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as tdist
from vican_b200 import dist as vdist

rank, world = vdist.init_process_group_from_env("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
NB = 800 << 20


def bus_id():
    p = torch.cuda.get_device_properties(local)
    return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)


def numa_of_gpu():
    try:
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bus_id()).read())
    except Exception as e:           # noqa: BLE001
        return "err:%s" % e


def measure(tag, buf, solo=False):
    dst = torch.empty(NB, dtype=torch.uint8, device=dev)
    out = []
    for it in range(4):
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not solo or rank == 0:
            a.record(); dst.copy_(buf, non_blocking=True); b.record(); torch.cuda.synchronize()
            out.append(NB / a.elapsed_time(b) / 1e6)
        if world > 1:
            tdist.barrier()
    bw = torch.tensor([max(out[1:]) if out else 0.0], device=dev)
    if world > 1:
        g = [torch.zeros_like(bw) for _ in range(world)]
        tdist.all_gather(g, bw)
    else:
        g = [bw]
    if rank == 0:
        print("%-28s GB/s per rank: %s  sum %.1f" % (tag, " ".join("%.1f" % x.item() for x in g), sum(x.item() for x in g)), flush=True)


if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            print("\n".join(l for l in txt.splitlines() if cmd[0] != "lscpu" or "NUMA" in l or "Model name" in l or l.startswith("CPU(s)")), flush=True)
        except Exception as e:       # noqa: BLE001
            print(cmd, "failed", e)
    try:
        print("cpuset.mems:", open("/sys/fs/cgroup/cpuset.mems.effective").read().strip(), "cpuset.cpus:", open("/sys/fs/cgroup/cpuset.cpus.effective").read().strip())
    except Exception as e:           # noqa: BLE001
        print("cgroup cpuset unreadable:", e)
info = [None] * world
me = (rank, bus_id(), numa_of_gpu(), len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4])
if world > 1:
    tdist.all_gather_object(info, me)
else:
    info = [me]
if rank == 0:
    for i in info:
        print("rank %d gpu %s numa %s affinity %d cpus %s.." % i, flush=True)

buf = torch.empty(NB, dtype=torch.uint8).pin_memory()
buf.fill_(1)
measure("default, solo rank 0", buf, solo=True)
measure("default, all ranks", buf)
del buf

from vican_b200 import hostmem      # noqa: E402
got = hostmem.bind_to_gpu_node(local)
res = [None] * world
if world > 1:
    tdist.all_gather_object(res, got)
else:
    res = [got]
if rank == 0:
    print("bind_to_gpu_node:", res, flush=True)
buf = torch.empty(NB, dtype=torch.uint8).pin_memory()
buf.fill_(1)
measure("bound, solo rank 0", buf, solo=True)
measure("bound, all ranks", buf)
if world > 1:
    tdist.destroy_process_group()
