#!/usr/bin/env python
"""Latency of the cross-rank sum of a camera accumulator (n_c x 9 doubles) on N GPUs of one box:
one-shot peer-memory kernel (csrc/peer.cuh) vs ncclAllReduce, CUDA-event timed back to back, plus
the stage stamps of the peer kernel.   torchrun --nproc-per-node N scripts/peer_probe.py [n_c]"""
import ctypes as C
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vican_b200 import _cabi, dist as vdist  # noqa: E402
from vican_b200.solver import _ptr, _stream  # noqa: E402


def main():
    n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
    rank, world = vdist.init_process_group_from_env("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    lib = _cabi.lib()
    comm = vdist.create_comm()
    assert comm.peer is not None
    for n in (9 * n_c, 3, 3 * n_c + 8, 9 * n_c, 3):
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        res = {}
        for name, fn in (("peer", lambda: lib.vb_peer_allreduce(comm.peer, _ptr(x), n, _stream())),
                         ("nccl", lambda: lib.vb_nccl_allreduce(comm.ctx, _ptr(x), n, _stream()))):
            for _ in range(20):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 200
            a.record()
            for _ in range(reps):
                fn()
                x.mul_(0.5)          # keep magnitudes bounded; also a dependent kernel between calls, as in the solver
            b.record(); torch.cuda.synchronize()
            res[name] = 1e3 * a.elapsed_time(b) / reps
        a.record()
        for _ in range(200):
            x.mul_(0.5)
        b.record(); torch.cuda.synchronize()
        base = 1e3 * a.elapsed_time(b) / 200
        lib.vb_peer_allreduce(comm.peer, _ptr(x), n, _stream())
        st = (C.c_uint64 * 4)()
        lib.vb_peer_stamps(comm.peer, st, _stream())
        if rank == 0:
            print("world %d, %7d doubles: peer %.1f us, nccl %.1f us per call (net of the %.1f us filler kernel); "
                  "peer stages: publish->flags %.1f, sums %.1f, tail %.1f us" %
                  (world, n, res["peer"] - base, res["nccl"] - base, base, 1e-3 * (st[1] - st[0]), 1e-3 * (st[2] - st[1]),
                   1e-3 * (st[3] - st[2])), flush=True)
    vdist.destroy_comm(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
