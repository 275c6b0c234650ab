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
def cpu_sample_solve(seed, n_c, n_t, cams_per_t, maxiter):
    """Oracle port (numpy/scipy restatement of vican/bipgo.py) on a bounded cfg4-shaped sample."""
    from oracle import vican_oracle as orc
    from vican_b200 import synthetic as syn
    g = syn.make_camera_network(seed, n_c, n_t, 1, cams_per_t, 1)
    marker_R = g.marker_R
    marker_t_inv0 = np.zeros((1, 3))
    t0 = time.perf_counter()
    orc.solve_arrays_oracle(g.cam, g.time, g.marker, g.R, g.t, g.w, 2.0 * g.w, marker_R, marker_t_inv0, 0,
                            n_c, n_t, maxiter, "conjugate_gradient")
    return time.perf_counter() - t0, g.n_edges


def cpu_baseline(full_edges, full_maxiter, sample=(4, 500, 10_000, 50, 4)):
    try:
        from threadpoolctl import threadpool_info
        threads = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        threads = os.cpu_count() or 1
    seed, n_c, n_t, d, it = sample
    sec, e_s = cpu_sample_solve(seed, n_c, n_t, d, it)
    iter_s_sample = it / sec
    # scaled to the metric's unit: iterations/s the CPU would deliver on the full graph if its
    # cost were linear in the edge count (optimistic for the CPU: its eigensolver and the
    # power-graph SpGEMM grow faster than E)
    value = iter_s_sample * e_s / full_edges
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle/vican_oracle.py (numpy/scipy port of vican/bipgo.py; /root/reference is absent on the "
                      "GPU box) on a cfg4-shaped sample: %d cameras, %d time nodes, %d edges, maxiter=%d: %.1f s "
                      "(%.3f iter/s on the sample), scaled linearly by edges to %d edges"
                      % (n_c, n_t, e_s, it, sec, iter_s_sample, full_edges),
            "sample_seconds": sec, "sample_iter_per_s": iter_s_sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seed, n_c, n_t, d, maxiter = WORKLOADS[args.workload]
    full_edges = n_t * d
    vals, secs = [], []
    sample = (4, 400, 8_000, 50, 3)
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue   # one warm-up pass is enough to page numpy/scipy in; keeps the arm within minutes
        cb = cpu_baseline(full_edges, maxiter, sample)
        if i >= args.warmup:
            vals.append(cb["value"]); secs.append(cb["sample_seconds"])
    cb["value"] = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "n_cameras": n_c, "n_time_nodes": n_t, "n_edges": full_edges,
                       "maxiter": maxiter, "lsqr_solver": "conjugate_gradient"},
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from vican_b200 import _cabi, dist as vdist, solver
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
    phases, launches, loop_ms = [], 0, []
    in_step = {"time": [0.0, 0], "cam": [0.0, 0]}   # CUDA-event time of the edge passes INSIDE the timed steps
    barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        res = one_solve(profile_events=True)
        phases.append(res.phase_ms)
        st = res.rot.stats
        in_step["time"][0] += st.time_pass_ms; in_step["time"][1] += st.time_pass_timed
        in_step["cam"][0] += st.cam_pass_ms; in_step["cam"][1] += st.cam_pass_timed
        launches += 13 + st.kernel_launches + 12 + 3 * (res.trans.iters + 8)
        loop_ms.append(res.phase_ms["rotation"])
    e1.record()
    torch.cuda.synchronize(); barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = maxiter * args.steps / (total_ms * 1e-3)
    st = res.rot.stats
    g = res.graph
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
                "frac": kern[dom]["gbs"] / peak, "traffic": 3.98e9 if "time" in dom else 4.01e9,
                "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch on cfg4 at 1 GPU "
                                  "(profiles/r1_edge_pass_history.md, capture prof_passes_r1e)",
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
        one_solve(host, to_host=True)
        k_e2e = max(1, min(args.steps, 3))
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            r2 = one_solve(host, to_host=True)
        torch.cuda.synchronize(); barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        d2h = sum(v.numel() * v.element_size() for v in (r2.Rw_c, r2.Rw_t, r2.x_c, r2.x_t))
        e2e = {"value": maxiter * k_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(sum_over_ranks(float(h2d))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(d2h))), "ms_per_step": 1e3 * e2e_s / k_e2e, "steps": k_e2e}
        del host

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(edges_total, maxiter)

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
            "eig_residual_rel": max(st.resid) / st.anorm if st.anorm else None, "cg_iters": res.trans.iters,
            "max_rot_err_vs_ground_truth_rad": gt_err,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cb,
        }
        emit(line)
    vdist.destroy_comm(comm)
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
