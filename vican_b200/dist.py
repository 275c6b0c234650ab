"""Edge-sharded multi-GPU plumbing: one process per GPU (torchrun), time nodes partitioned into
contiguous ranges, camera-side quantities replicated and summed with one NCCL all-reduce per
camera pass.  torch.distributed is used only to bootstrap (broadcast the NCCL
unique id) and for barriers / timing reductions in the benchmark; the data-path collective is
issued by the extension itself on the solver's stream."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _cabi
from .solver import Comm


def init_process_group_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def create_comm(peer_capacity_doubles: int = 9 * 16384 + 64) -> Optional[Comm]:
    """Communicator of the extension for the current process group (None on 1 rank): an NCCL
    communicator plus, when every rank can map its peers (one box, <= 8 GPUs, CUDA IPC), the
    NVLink peer-memory windows of csrc/peer.cuh sized for ``peer_capacity_doubles`` per exchange
    (default: the camera accumulator of 16 k cameras).  ``VICAN_B200_COLLECTIVE=nccl`` disables
    the windows."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    lib = _cabi.lib()
    if not lib.vb_nccl_available():
        raise RuntimeError("libnccl.so.2 could not be resolved by the extension")
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _cabi.check(lib.vb_nccl_unique_id(buf, 128), "vb_nccl_unique_id")
    obj = [bytes(buf.raw)]
    dist.broadcast_object_list(obj, src=0)
    idbuf = C.create_string_buffer(obj[0], 128)
    ctx = C.c_void_p()
    _cabi.check(lib.vb_nccl_init(idbuf, 128, rank, world, C.byref(ctx)), "vb_nccl_init")
    comm = Comm(ctx.value, rank, world)
    # NVLink peer-memory windows for the one-shot / fused all-reduce (csrc/peer.cuh).  All ranks must
    # agree: the windows are used only if EVERY rank could map every other rank's window.
    mode = os.environ.get("VICAN_B200_COLLECTIVE", "peer").lower()
    if mode not in ("peer", "nccl"):
        raise ValueError("VICAN_B200_COLLECTIVE must be 'peer' or 'nccl'")
    if mode == "peer" and world <= 8:
        cap = int(peer_capacity_doubles)
        pctx, handle = C.c_void_p(), C.create_string_buffer(64)
        ok = lib.vb_peer_create(rank, world, cap, C.byref(pctx), handle) == 0
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw) if ok else None)
        if ok and all(h is not None for h in handles):
            ok = lib.vb_peer_connect(pctx, C.create_string_buffer(b"".join(handles), 64 * world)) == 0
        else:
            ok = False
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))          # also the barrier between connect and first use
        if all(flags):
            comm.peer, comm.peer_capacity = pctx.value, cap
        elif pctx.value:
            lib.vb_peer_destroy(pctx)
    return comm


def destroy_comm(comm: Optional[Comm]):
    if comm is not None:
        import torch.distributed as dist
        lib = _cabi.load_library()
        if comm.peer is not None:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()                            # nobody may still read a window that is about to be freed
            lib.vb_peer_destroy(comm.peer)
            comm.peer = None
        lib.vb_nccl_destroy(comm.ctx)


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [lo, hi) of ``n_items`` owned by ``rank`` (balanced to +-1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
