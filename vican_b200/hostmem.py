"""Placement of a rank's host buffers next to its GPU.

With one process per GPU (torchrun) every rank uploads its shard of the detections at the same time.  A pinned
buffer that sits on the other socket crosses the inter-socket link on its way to the GPU, and with 4-8 ranks
uploading at once that link -- not PCIe -- bounds the upload.  ``bind_to_gpu_node`` moves the calling process onto
the CPUs of the GPU's NUMA node (when the cpuset allows) and makes that node the preferred one for its page
allocations, so pinned buffers allocated afterwards land there.  Nothing here touches the device."""
from __future__ import annotations

import ctypes
import os
import platform
from typing import Optional

_MPOL_PREFERRED = 1
_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}


def _parse_cpulist(txt: str) -> set:
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(local_gpu: int) -> Optional[int]:
    """NUMA node of a CUDA device from sysfs (None when the platform does not say: VMs report -1)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_gpu)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        return node if node >= 0 else None
    except Exception:                # noqa: BLE001 -- placement is an optimisation, never an error
        return None


def bind_to_gpu_node(local_gpu: int) -> dict:
    """Prefer the GPU's NUMA node for this process's CPUs and page allocations.  Returns what was done:
    ``{"node": n | None, "cpus": number of CPUs kept | None, "mempolicy": bool}``.  ``VICAN_B200_NUMA=0`` disables it."""
    done = {"node": None, "cpus": None, "mempolicy": False}
    if os.environ.get("VICAN_B200_NUMA", "1") == "0":
        return done
    node = gpu_numa_node(local_gpu)
    if node is None:
        return done
    done["node"] = node
    try:
        local_cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        keep = local_cpus & os.sched_getaffinity(0)
        if keep:
            os.sched_setaffinity(0, keep)
            done["cpus"] = len(keep)
    except Exception:                # noqa: BLE001
        pass
    nr = _SYS_SET_MEMPOLICY.get(platform.machine())
    if nr is not None and node < 1024:
        try:
            mask = (ctypes.c_ulong * 16)()
            bits = 8 * ctypes.sizeof(ctypes.c_ulong)
            mask[node // bits] = 1 << (node % bits)
            libc = ctypes.CDLL(None, use_errno=True)
            libc.syscall.restype = ctypes.c_long
            rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(_MPOL_PREFERRED), mask, ctypes.c_ulong(16 * bits + 1))
            done["mempolicy"] = rc == 0
        except Exception:            # noqa: BLE001
            pass
    return done
