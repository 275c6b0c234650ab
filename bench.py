#!/usr/bin/env python
"""Benchmark of the hot path: bipartite SE(3) synchronisation on a cfg4-shaped synthetic camera
network (BASELINE.json configs[3]: 10 k cameras, 1 M marker-timestep nodes, 50 M edges), the
configuration the metric "PGO solve ms & primal-dual iter/s at 1/2/4/8 B200; HBM GB/s vs peak"
is quoted on.  It fits one GPU, so it is also the N=1 workload; at N>1 the SAME graph is
edge-sharded by time-node range (strong scaling) with one NCCL all-reduce per camera pass.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
    python bench.py --impl reference ...                      (CPU arm: numpy/scipy oracle port)

A "step" is one full solve: ingestion of the device-resident raw detections, `maxiter`
primal-dual iterations, translation CG.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# rank 0 prints exactly ONE line on stdout.  Libraries write banners there (NCCL prints its version
# at NCCL_DEBUG=VERSION and above), so file descriptor 1 is pointed at stderr for the whole run
# and the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (seed, n_c, n_t, cams_per_t, maxiter)
    "cfg4": (4, 10_000, 1_000_000, 50, 10),
    "cfg4_small": (4, 2_000, 100_000, 50, 10),      # quick functional runs
}
METRIC = "primal_dual_iter_per_s"
UNIT = "iter/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    """Threads the numpy / scipy path can use (BLAS pool); the Python loops and scipy's sparse kernels
    around them are single-threaded."""
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_solve(seed, n_c, n_t, cams_per_t, maxiter):
    """Oracle port (numpy/scipy restatement of vican/bipgo.py) on a bounded cfg4-shaped sample."""
    from oracle import vican_oracle as orc
    from vican_b200 import synthetic as syn
    g = syn.make_camera_network(seed, n_c, n_t, 1, cams_per_t, 1)
    marker_R = g.marker_R
    marker_t_inv0 = np.zeros((1, 3))
    iters = []
    t0 = time.perf_counter()
    orc.solve_arrays_oracle(g.cam, g.time, g.marker, g.R, g.t, g.w, 2.0 * g.w, marker_R, marker_t_inv0, 0,
                            n_c, n_t, maxiter, "conjugate_gradient", timings=iters)
    return time.perf_counter() - t0, g.n_edges, iters


def cpu_baseline(full_edges, full_maxiter, sample=(4, 500, 10_000, 50, 4)):
    """The oracle port timed on a bounded cfg4-shaped sample.  The sample runs `it` <= maxiter primal-dual
    iterations; the solve at the FULL maxiter on the sample is the measured set-up + translation time plus
    maxiter x the measured mean iteration time.  `value` = the metric (primal-dual iterations / s of a whole
    solve) the CPU would deliver on the full graph if its cost were linear in the edge count -- optimistic for
    the CPU: its dense eigen-solve grows with n_c^3 and the power-graph SpGEMM with E * degree."""
    seed, n_c, n_t, d, it = sample
    sec_run, e_s, iters = cpu_sample_solve(seed, n_c, n_t, d, it)
    per_iter = float(np.mean(iters))
    sec = sec_run - float(np.sum(iters)) + full_maxiter * per_iter
    iter_s_sample = full_maxiter / sec
    value = iter_s_sample * e_s / full_edges
    return {"value": value, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
            "sample": "oracle/vican_oracle.py (numpy/scipy port of vican/bipgo.py; /root/reference is absent on the "
                      "GPU box) on a cfg4-shaped sample: %d cameras, %d time nodes, %d edges; measured %d iterations in "
                      "%.1f s (%.1f s per iteration + %.1f s set-up and translation CG) = %.1f s at maxiter=%d "
                      "(%.4f iter/s on the sample), scaled linearly by edges (x%.0f) to %d edges; BLAS pool of %d "
                      "threads, everything else single-threaded"
                      % (n_c, n_t, e_s, it, sec_run, per_iter, sec_run - float(np.sum(iters)), sec, full_maxiter,
                         iter_s_sample, full_edges / e_s, full_edges, cpu_threads()),
            "sample_seconds": sec_run, "sample_iterations": it, "sample_seconds_per_iteration": per_iter,
            "sample_seconds_at_full_maxiter": sec, "sample_iter_per_s": iter_s_sample, "sample_edges": e_s,
            "edge_scale_factor": full_edges / e_s}


REFERENCE_SAMPLE = (4, 1000, 100_000, 50, 2)     # cfg4 at 1/10 scale: 5 M edges, two primal-dual iterations


def run_reference(args):
    """CPU arm: the reference's own algorithm (oracle port; the reference is pure Python and cannot travel
    to the GPU box) on cfg4 at 1/10 scale -- 1 000 cameras, 100 000 time nodes, 5 M edges, the size
    SURVEY.md 8d prescribes for the CPU run -- for TWO primal-dual iterations plus the translation CG, once
    (about two minutes of CPU work; `--steps` / `--warmup` bound the GPU arm, this arm always runs one warm-up
    on a tiny graph and one timed sample).  The solve at maxiter = 10 on the sample is assembled from the measured
    set-up, per-iteration and translation times, and its rate is scaled by the edge ratio (x10) to cfg4; both
    factors are stated in the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seed, n_c, n_t, d, maxiter = WORKLOADS[args.workload]
    full_edges = n_t * d
    cpu_sample_solve(4, 50, 400, 10, 1)          # pages numpy / scipy in
    sample = REFERENCE_SAMPLE if args.workload == "cfg4" else (4, 200, 4_000, 50, 2)
    if os.environ.get("VICAN_B200_REF_SAMPLE") == "small":      # tests/test_bench_contract.py: contract only
        sample = (4, 200, 4_000, 50, 2)
    cb = cpu_baseline(full_edges, maxiter, sample)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "steps_run": 1, "ms_per_step": 1e3 * cb["sample_seconds"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_cameras": n_c, "n_time_nodes": n_t, "n_edges": full_edges,
                       "maxiter": maxiter, "lsqr_solver": "conjugate_gradient",
                       "reference_sample": {"n_cameras": sample[1], "n_time_nodes": sample[2], "n_edges": cb["sample_edges"],
                                            "maxiter": sample[4], "edge_scale_factor": cb["edge_scale_factor"],
                                            "note": "the reference cannot run cfg4 itself (dense 30 000^2 power graph); "
                                                    "value = sample iter/s divided by the edge ratio"}},
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------- BASELINE configs 0-2, 4 (dict drop-in)
def configs_block(names):
    """BASELINE.json configs[0, 1, 2, 4] at FULL size through the dictionary drop-in
    (`bipartite_se3sync` / `object_bipartite_se3sync`: flatten on the host, H2D, solve, D2H), next to the
    oracle port's wall time on the SAME dictionaries (like for like, same host) and the parity of the two
    results.  cfg5 (maxiter 500) is timed on the device only: its oracle run takes ~12 min (parity of that
    config: tests/test_gpu_fullsize.py at 10 % of the time nodes)."""
    import torch
    from oracle import vican_oracle as orc
    from vican_b200 import bipgo, synthetic as syn
    from vican_b200.geometry import SE3
    nr, nt, ef = syn.default_callables()
    g0 = syn.make_camera_network(0, 6, 20, 3, 3, 2)
    e0, c0 = syn.to_edge_dict(g0, SE3)
    bipgo.bipartite_se3sync(e0, c0, nr, nt, ef, 2, "conjugate_gradient")     # warm the library
    out_block = {}
    for name in names:
        g, p = syn.make_config(name)
        edges, cons = syn.to_edge_dict(g, SE3)
        best = None
        for _ in range(2 if g.n_edges < 500_000 else 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if g.kind == "object":
                out = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **p)
            else:
                out = bipgo.bipartite_se3sync(edges, cons, nr, nt, ef, dtype=np.float64, **p)
            wall = time.perf_counter() - t0
            if best is None or wall < best[0]:
                best = (wall, dict(bipgo.last_info))
        wall, info = best
        rec = {"api": "object_bipartite_se3sync" if g.kind == "object" else "bipartite_se3sync",
               "detections": g.n_edges, "n_c": info["n_c"], "n_t": info["n_t"], "n_edges": info["n_edges"],
               "maxiter": p["maxiter"], "lsqr_solver": p["lsqr_solver"],
               "wall_ms": 1e3 * wall, "device_ms": 1e3 * info["device_seconds"],
               "host_flatten_ms": 1e3 * (wall - info["device_seconds"]),
               "iter_per_s_wall": p["maxiter"] / wall, "iter_per_s_device": p["maxiter"] / info["device_seconds"],
               "trans_iters": info["trans_iters"], "inner_per_outer": info["inner_per_outer"][:10]}
        if p["maxiter"] <= 20:
            t0 = time.perf_counter()
            if g.kind == "object":
                ref = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, se3_cls=SE3, **p)
            else:
                ref = orc.bipartite_se3sync_oracle(edges, cons, nr, nt, ef, **p)
            rec["cpu_port_wall_ms"] = 1e3 * (time.perf_counter() - t0)
            rec["speedup_wall"] = rec["cpu_port_wall_ms"] / rec["wall_ms"]
            keys = sorted(ref.keys())
            Ra = np.stack([np.asarray(out[k].R(), np.float64) for k in keys]); Rb = np.stack([ref[k][0] for k in keys])
            ta = np.stack([out[k].t() for k in keys]); tb = np.stack([ref[k][1] for k in keys])
            D = np.transpose(Ra, (0, 2, 1)) @ Rb
            sk = 0.5 * np.sqrt((D[:, 2, 1] - D[:, 1, 2]) ** 2 + (D[:, 0, 2] - D[:, 2, 0]) ** 2 + (D[:, 1, 0] - D[:, 0, 1]) ** 2)
            rec["rot_err_rad"] = float(np.arctan2(sk, 0.5 * (np.trace(D, axis1=1, axis2=2) - 1.0)).max())
            rec["rel_t_err"] = float((np.linalg.norm(ta - tb, axis=1) / np.maximum(np.linalg.norm(tb, axis=1), 1e-300)).max())
        else:
            rec["cpu_port_wall_ms"] = None
            rec["note"] = "oracle not run at maxiter=%d (minutes); parity: tests/test_gpu_fullsize.py" % p["maxiter"]
        out_block[name] = rec
        del edges, cons, out
    return out_block


# ------------------------------------------------------------------------------ GPU arm
def multi_gpu_check(rank, world, dev, comm, solver, vdist, make_scaled_network, tdist,
                    shape=(7, 2_000, 100_000, 50, 6)):
    """Edge-sharded solve of a 5 M-edge cfg4-shaped graph on all ranks vs the single-GPU solve of the same
    graph on rank 0: camera results bitwise equal across ranks, everything within 1e-9 rad / 1e-8 (relative
    translation) of the 1-GPU result, same CG iteration count.  Raises on failure."""
    import torch
    seed, n_c, n_t, d, maxiter = shape
    lo, hi = vdist.shard_range(n_t, rank, world)
    det = make_scaled_network(seed, n_c, n_t, d, lo, hi, block=5_000, device=dev)
    I9 = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
    q0 = torch.zeros((1, 3), dtype=torch.float64, device=dev)
    res = solver.solve_arrays(det.cam, det.time, det.marker, det.R, det.t, det.k_r, det.k_t, I9, q0, n_c, det.n_t,
                              maxiter, "conjugate_gradient", comm=comm)
    # replicated camera state: identical bits on every rank
    ref_c = torch.cat([res.Rw_c.reshape(-1), res.x_c.reshape(-1)]).clone()
    tdist.broadcast(ref_c, src=0)
    same = torch.tensor([1.0 if torch.equal(ref_c, torch.cat([res.Rw_c.reshape(-1), res.x_c.reshape(-1)])) else 0.0],
                        dtype=torch.float64, device=dev)
    tdist.all_reduce(same, op=tdist.ReduceOp.MIN)
    sizes = [vdist.shard_range(n_t, r, world) for r in range(world)]
    parts_R = [torch.empty((b - a, 3, 3), dtype=torch.float64, device=dev) for a, b in sizes]
    parts_x = [torch.empty((b - a, 3), dtype=torch.float64, device=dev) for a, b in sizes]
    tdist.all_gather(parts_R, res.Rw_t.contiguous())
    tdist.all_gather(parts_x, res.x_t.contiguous())
    out = None
    if rank == 0:
        full = make_scaled_network(seed, n_c, n_t, d, 0, n_t, block=5_000, device=dev)
        one = solver.solve_arrays(full.cam, full.time, full.marker, full.R, full.t, full.k_r, full.k_t, I9, q0, n_c, n_t,
                                  maxiter, "conjugate_gradient")

        def geo(A, B):
            D = A.transpose(1, 2) @ B
            sk = 0.5 * torch.sqrt((D[:, 2, 1] - D[:, 1, 2]) ** 2 + (D[:, 0, 2] - D[:, 2, 0]) ** 2 + (D[:, 1, 0] - D[:, 0, 1]) ** 2)
            return float(torch.atan2(sk, 0.5 * (D.diagonal(dim1=1, dim2=2).sum(1) - 1.0)).max())

        def relt(a, b):
            return float(((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-300)).max())
        out = {"graph": {"n_cameras": n_c, "n_time_nodes": n_t, "n_edges": n_t * d, "maxiter": maxiter},
               "camera_state_bitwise_equal_across_ranks": bool(same.item() == 1.0),
               "rot_err_cam_rad": geo(res.Rw_c, one.Rw_c), "rot_err_time_rad": geo(torch.cat(parts_R), one.Rw_t),
               "rel_t_err_cam": relt(res.x_c, one.x_c), "rel_t_err_time": relt(torch.cat(parts_x), one.x_t),
               "cg_iters": [res.trans.iters, one.trans.iters], "tolerance": {"rot_rad": 1e-9, "rel_t": 1e-8}}
        out["ok"] = bool(out["camera_state_bitwise_equal_across_ranks"] and out["cg_iters"][0] == out["cg_iters"][1]
                         and max(out["rot_err_cam_rad"], out["rot_err_time_rad"]) <= 1e-9
                         and max(out["rel_t_err_cam"], out["rel_t_err_time"]) <= 1e-8)
        if not out["ok"]:
            raise SystemExit("bench.py: multi-GPU consistency check FAILED: %s" % json.dumps(out))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg5",
                    help="BASELINE configs measured through the dict drop-in next to the CPU port (N=1 only); '' = none")
    ap.add_argument("--no-multi-gpu-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from vican_b200 import _cabi, dist as vdist, hostmem, solver
    from vican_b200.synthetic_device import make_scaled_network
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    rank, world = vdist.init_process_group_from_env("nccl")
    if args.gpus != world:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d; for N>1 launch with "
                         "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N`" % (args.gpus, world))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # N ranks upload their shards at once (e2e arm): keep every rank's pinned buffers on its GPU's NUMA node
    placement = hostmem.bind_to_gpu_node(local) if world > 1 else None
    lib = _cabi.lib()
    comm = vdist.create_comm()
    import torch.distributed as tdist

    def barrier():
        if world > 1:
            tdist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return float(t.item())

    seed, n_c, n_t_glob, d, maxiter = WORKLOADS[args.workload]
    lo, hi = vdist.shard_range(n_t_glob, rank, world)
    det = make_scaled_network(seed, n_c, n_t_glob, d, lo, hi, device=dev)
    torch.cuda.synchronize()
    markerC = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
    marker_q = torch.zeros((1, 3), dtype=torch.float64, device=dev)

    def one_solve(src=det, to_host=False, profile_events=False):
        return solver.solve_arrays(src.cam, src.time, src.marker, src.R, src.t, src.k_r, src.k_t, markerC, marker_q,
                                   n_c, src.n_t, maxiter, "conjugate_gradient", comm=comm, to_host=to_host,
                                   profile_events=profile_events, reuse_host_buffers=to_host)

    # ---- device-resident timing: W warm-up, K timed steps, barrier + synchronize on both sides
    res = None
    for _ in range(args.warmup):
        res = one_solve()
    barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases, loop_ms = [], []
    launches0 = int(lib.vb_launch_count())
    in_step = {"time": [0.0, 0], "cam": [0.0, 0]}   # CUDA-event time of the edge passes INSIDE the timed steps
    barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        res = one_solve(profile_events=True)
        phases.append(res.phase_ms)
        st = res.rot.stats
        in_step["time"][0] += st.time_pass_ms; in_step["time"][1] += st.time_pass_timed
        in_step["cam"][0] += st.cam_pass_ms; in_step["cam"][1] += st.cam_pass_timed
        loop_ms.append(res.phase_ms["rotation"])
    e1.record()
    torch.cuda.synchronize(); barrier()
    launches = int(lib.vb_launch_count()) - launches0      # this library's kernels executed inside the timed region
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = maxiter * args.steps / (total_ms * 1e-3)
    st = res.rot.stats
    g = res.graph
    cg_iters = res.trans.iters
    edges_total = int(sum_over_ranks(float(g.n_edges)))

    # ---- sanity: the solve recovers the synthetic ground truth up to gauge (noise-level error)
    Rgt = det.gt_cam_R
    G = Rgt[0] @ res.Rw_c[0].T
    rel = G @ res.Rw_c
    cos = ((rel * Rgt).sum(dim=(1, 2)) - 1.0) * 0.5
    gt_err = float(torch.acos(cos.clamp(-1, 1)).max().item())

    # ---- per-kernel roofline: the two edge passes, timed alone with CUDA events (inputs 3.9 GB >> L2)
    import ctypes as C
    gs = int(lib.vb_gather_stride())                                   # padded gather layout (3 rows x 4, one 128-byte line)
    X = torch.randn((n_c, gs), dtype=torch.float64, device=dev)
    lamT = torch.randn((g.n_t, 9), dtype=torch.float64, device=dev)
    Wt = torch.zeros((g.n_t, gs), dtype=torch.float64, device=dev)
    Y = torch.zeros((n_c, 9), dtype=torch.float64, device=dev)
    ptr, stream = solver._ptr, solver._stream
    kern = {}
    for name, fn in (("edge_pass_kernel<0> (time pass)", lambda: lib.vb_pass_time(C.byref(g.cgraph), 0, ptr(X), ptr(lamT), ptr(Wt), stream())),
                     ("edge_pass_kernel<3> + tile_combine (camera pass)", lambda: lib.vb_pass_cam(C.byref(g.cgraph), ptr(Wt), ptr(Y), stream()))):
        for _ in range(3):
            fn()
        reps = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        nbytes = g.pass_bytes("time" if "time" in name else "cam")
        kern[name] = {"ms": ms, "bytes": nbytes, "gbs": nbytes / ms * 1e-6}
    peak, peak_src = measured_peaks()
    n_time, n_cam = st.time_passes, st.cam_passes
    # the roofline entry uses the duration measured INSIDE the timed steps (events around every executed
    # launch on the solver's stream); the isolated micro-loop above is reported next to it
    for name in kern:
        tot_ms, cnt = in_step["time" if "time" in name else "cam"]
        kern[name]["isolated_ms"] = kern[name]["ms"]
        kern[name]["isolated_gbs"] = kern[name]["gbs"]
        if cnt > 0:
            kern[name]["ms"] = tot_ms / cnt
            kern[name]["gbs"] = kern[name]["bytes"] / kern[name]["ms"] * 1e-6
        kern[name]["timed_launches"] = cnt
    share = {k: (n_time if "time" in k else n_cam) * v["ms"] / phases[-1]["rotation"] for k, v in kern.items()}
    dom = max(kern, key=lambda k: share[k])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbs"] / peak, "traffic": 4.060e9 if "time" in dom else 4.017e9,
                "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch on cfg4 at 1 GPU "
                                  "(profiles/r2_edge_pass.md, capture prof_passes_r2b)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kern[dom]["bytes"], "ms_per_launch": kern[dom]["ms"],
                "measured": "CUDA events around every executed launch inside the timed steps",
                "share_of_rotation_stage": share[dom], "frac_of_nominal_8000": kern[dom]["gbs"] / 8000.0,
                "kernels": {k: dict(v, frac=v["gbs"] / peak, isolated_frac=v["isolated_gbs"] / peak,
                                    launches_per_step=(n_time if "time" in k else n_cam),
                                    share_of_rotation_stage=share[k]) for k, v in kern.items()}}
    if world > 1:
        roofline["traffic"] = None
    del X, lamT, Wt, Y

    # ---- end to end through the public array API with HOST buffers (pinned): H2D + solve + D2H
    e2e = None
    if not args.no_e2e:
        import dataclasses
        host = dataclasses.replace(det, **{f: getattr(det, f).cpu().pin_memory()
                                           for f in ("cam", "time", "marker", "R", "t", "k_r", "k_t")})
        h2d = sum(getattr(host, f).numel() * getattr(host, f).element_size()
                  for f in ("cam", "time", "marker", "R", "t", "k_r", "k_t"))
        barrier()          # pinning the host buffers takes seconds and not the same time on every rank
        res = g = None            # the device-resident result (graph: ~10 GB) is no longer needed
        one_solve(host, to_host=True)
        one_solve(host, to_host=True)
        k_e2e = max(1, min(args.steps, 10))
        e2e_phases = []
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            r2 = one_solve(host, to_host=True)
            e2e_phases.append(r2.phase_ms)
        torch.cuda.synchronize(); barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        d2h = sum(v.numel() * v.element_size() for v in (r2.Rw_c, r2.Rw_t, r2.x_c, r2.x_t))
        e2e = {"value": maxiter * k_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(sum_over_ranks(float(h2d))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(d2h))), "ms_per_step": 1e3 * e2e_s / k_e2e, "steps": k_e2e,
               "api": "vican_b200.solver.solve_arrays (pinned host arrays in, pinned host results out)",
               "phase_ms": {k: float(np.mean([ph[k] for ph in e2e_phases])) for k in e2e_phases[0]}}
        del host, r2

    # ---- N > 1: the sharded solve must be the SAME solve (replicated state bitwise equal on every rank,
    # ---- results equal to a 1-GPU solve of the same graph) -- checked on a 5 M-edge graph of the same shape
    mg_check = None
    if world > 1 and not args.no_multi_gpu_check:
        res = g = None
        mg_check = multi_gpu_check(rank, world, dev, comm, solver, vdist, make_scaled_network, tdist)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(edges_total, maxiter, (4, 500, 10_000, 50, 3))

    # ---- BASELINE configs 0-2 and 4 through the dictionary drop-in, like for like with the CPU port
    cfg_block = None
    if rank == 0 and world == 1 and args.configs:
        res = g = None
        del det
        torch.cuda.empty_cache()
        cfg_block = configs_block([c for c in args.configs.split(",") if c])

    if rank == 0:
        mean = lambda k: float(np.mean([p[k] for p in phases]))  # noqa: E731
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_cameras": n_c, "n_time_nodes": n_t_glob, "n_edges": edges_total,
                       "cams_per_node": d, "maxiter": maxiter, "lsqr_solver": "conjugate_gradient",
                       "parallelism": "edge-sharded by time-node range x%d" % world,
                       "collective": (None if world == 1 else
                                      "one-shot all-reduce over NVLink peer memory, fused into the camera pass"
                                      if comm.peer is not None else "ncclAllReduce per camera pass"),
                       "l2_policy": "inputs (%.1f GB of edge blocks per pass) larger than L2" % (76e-9 * edges_total / world)},
            "solve_ms": ms_per_step, "loop_iter_per_s": maxiter / (float(np.mean(loop_ms)) * 1e-3),
            "phase_ms": {"ingest": mean("ingest"), "rotation": mean("rotation"), "translation": mean("translation")},
            "passes_per_step": {"time": st.time_passes, "cam": st.cam_passes, "lobpcg_steps": st.lobpcg_steps,
                                "inner_per_outer": list(st.inner_per_outer[:maxiter])},
            "eig_residual_rel": max(st.resid) / st.anorm if st.anorm else None, "cg_iters": cg_iters,
            "max_rot_err_vs_ground_truth_rad": gt_err,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "gpu_launches_source": "vb_launch_count(): the library's own tally of executed vb:: kernels (CUB excluded)",
            "roofline": roofline, "cpu_baseline": cb, "multi_gpu_check": mg_check, "configs": cfg_block,
            "host_placement": placement,
        }
        emit(line)
    vdist.destroy_comm(comm)
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
