"""Worker for the multi-rank tests (launched with torch.distributed.run).

  mode "gpu" : every rank solves its time-node shard of one synthetic graph with the NCCL hook;
               rank 0 also solves the whole graph alone and compares (edge-sharded == single GPU).
  mode "cpu" : gloo, no CUDA: checks the host-side sharding logic with the numpy model of the
               device algorithm -- per-shard camera passes summed by all_reduce equal the
               unsharded pass, shard ranges tile the node set, the NCCL-id broadcast path works.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main_cpu():
    import torch
    import torch.distributed as dist
    from oracle import device_model as dm, vican_oracle as orc
    from vican_b200 import synthetic as syn
    from vican_b200.dist import init_process_group_from_env, shard_range
    rank, world = init_process_group_from_env("gloo")
    assert world == 2
    g = syn.make_camera_network(3, 15, 90, 4, 5, 2)
    blk = orc.fold_blocks(g.R, g.w, g.marker, g.marker_R, 0)
    pc, pt, B, a = orc.aggregate_pairs(g.cam, g.time, blk, g.w, g.n_times)
    n_c, n_t = g.n_cams, g.n_times
    ranges = [shard_range(n_t, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_t and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    lo, hi = ranges[rank]
    sel = (pt >= lo) & (pt < hi)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n_c, 3, 3))
    lamT = rng.standard_normal((n_t, 3, 3))
    # local time pass on the owned nodes, local camera pass, all-reduce of the camera accumulator
    Z = dm.pass_time(pc[sel], pt[sel] - lo, B[sel], X, hi - lo)
    W = lamT[lo:hi] @ Z
    Y = dm.pass_cam(pc[sel], pt[sel] - lo, B[sel], W, n_c)
    Yt = torch.from_numpy(Y.copy())
    dist.all_reduce(Yt)
    Y_full = dm.pass_cam(pc, pt, B, lamT @ dm.pass_time(pc, pt, B, X, n_t), n_c)
    assert np.abs(Yt.numpy() - Y_full).max() < 1e-12 * np.abs(Y_full).max()
    # degrees: local camera degree sums to the global one
    deg = np.zeros(n_c); np.add.at(deg, pc[sel], a[sel])
    dt = torch.from_numpy(deg); dist.all_reduce(dt)
    deg_full = np.zeros(n_c); np.add.at(deg_full, pc, a)
    assert np.abs(dt.numpy() - deg_full).max() < 1e-12
    # id broadcast path used for the NCCL bootstrap
    obj = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK")
    dist.destroy_process_group()


def main_gpu():
    import torch
    import torch.distributed as dist
    from vican_b200 import dist as vdist, solver
    from vican_b200.synthetic_device import make_scaled_network
    rank, world = vdist.init_process_group_from_env("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = vdist.create_comm()
    want_peer = os.environ.get("VICAN_B200_COLLECTIVE", "peer") == "peer"
    assert (comm.peer is not None) == want_peer, "peer windows expected on a single NVLink box"
    if comm.peer is not None:
        # one-shot peer-memory all-reduce (csrc/peer.cuh) against NCCL, 30 back-to-back calls of mixed
        # sizes (both buffer parities, odd lengths, the full capacity); result bitwise equal on all ranks
        from vican_b200 import _cabi
        lib = _cabi.lib()
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        sizes = [1, 7, 1000, 9 * 600, 3 * 600 + 8, 12345, comm.peer_capacity] * 4 + [3, 2]
        for i, n in enumerate(sizes):
            x = torch.randn(n, dtype=torch.float64, device=dev, generator=gen)
            ref = x.clone()
            dist.all_reduce(ref)
            comm.allreduce(lib, x)
            assert (x - ref).abs().max().item() <= 1e-13 * max(1.0, ref.abs().max().item()), (i, n)
            same = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(same, x)
            assert all(torch.equal(same[0], s) for s in same), "peer all-reduce must be bitwise identical on all ranks"
        if rank == 0:
            print("PEER_ALLREDUCE_OK")
    seed, n_c, n_t, d, maxiter = 7, 600, 24_000, 30, 5
    lo, hi = vdist.shard_range(n_t, rank, world)
    det = make_scaled_network(seed, n_c, n_t, d, lo, hi, block=4000, device=dev)
    I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
    q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
    res = solver.solve_arrays(det.cam, det.time, det.marker, det.R, det.t, det.k_r, det.k_t, I9, q0, n_c, det.n_t,
                              maxiter, "conjugate_gradient", comm=comm)
    # gather the time-node results of every rank on rank 0
    parts_R = [torch.empty((vdist.shard_range(n_t, r, world)[1] - vdist.shard_range(n_t, r, world)[0], 3, 3),
                           dtype=torch.float64, device=dev) for r in range(world)]
    parts_x = [torch.empty((p.shape[0], 3), dtype=torch.float64, device=dev) for p in parts_R]
    dist.all_gather(parts_R, res.Rw_t.contiguous())
    dist.all_gather(parts_x, res.x_t.contiguous())
    ok = True
    if rank == 0:
        full = make_scaled_network(seed, n_c, n_t, d, 0, n_t, block=4000, device=dev)
        ref = solver.solve_arrays(full.cam, full.time, full.marker, full.R, full.t, full.k_r, full.k_t, I9, q0, n_c,
                                  n_t, maxiter, "conjugate_gradient")
        from util import geodesic_rad, rel_translation_err
        ec = geodesic_rad(res.Rw_c.cpu().numpy(), ref.Rw_c.cpu().numpy()).max()
        et = geodesic_rad(torch.cat(parts_R).cpu().numpy(), ref.Rw_t.cpu().numpy()).max()
        xc = rel_translation_err(res.x_c.cpu().numpy(), ref.x_c.cpu().numpy()).max()
        xt = rel_translation_err(torch.cat(parts_x).cpu().numpy(), ref.x_t.cpu().numpy()).max()
        print("DIST_GPU world=%d collective=%s rot_c=%.2e rot_t=%.2e x_c=%.2e x_t=%.2e cg_iters=%d/%d" %
              (world, "peer(fused)" if comm.peer is not None else "nccl", ec, et, xc, xt, res.trans.iters, ref.trans.iters))
        ok = ec < 1e-9 and et < 1e-9 and xc < 1e-8 and xt < 1e-8 and res.trans.iters == ref.trans.iters
    # lsqr_solver="direct" (LSQR replay) on the same shards: three collectives per bidiagonalisation step
    tr_d = solver.solve_translations(res.graph, res.rot, det.t, q0, "direct", comm=comm)
    parts_d = [torch.empty((p.shape[0], 3), dtype=torch.float64, device=dev) for p in parts_R]
    dist.all_gather(parts_d, tr_d.x_t.contiguous())
    if rank == 0:
        ref_d = solver.solve_translations(ref.graph, ref.rot, full.t, q0, "direct")
        dc = rel_translation_err(tr_d.x_c.cpu().numpy(), ref_d.x_c.cpu().numpy()).max()
        dt = rel_translation_err(torch.cat(parts_d).cpu().numpy(), ref_d.x_t.cpu().numpy()).max()
        print("DIST_GPU direct: x_c=%.2e x_t=%.2e itn=%d/%d istop=%d/%d" % (dc, dt, tr_d.iters, ref_d.iters, tr_d.istop, ref_d.istop))
        ok = ok and dc < 1e-6 and dt < 1e-6 and tr_d.iters == ref_d.iters and tr_d.istop == ref_d.istop
        print("DIST_GPU_OK" if ok else "DIST_GPU_FAIL")
    vdist.destroy_comm(comm)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    (main_gpu if sys.argv[1] == "gpu" else main_cpu)()
