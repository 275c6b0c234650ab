"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the
C ABI, against the CPU oracle / the committed reference goldens on identical inputs.

Tolerances are BASELINE.json's: rotation geodesic error <= 1e-6 rad, relative translation
error <= 1e-6 per node (fp64)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import device_model as dm            # noqa: E402
from oracle import vican_oracle as orc           # noqa: E402
from vican_b200 import synthetic as syn          # noqa: E402
from vican_b200.geometry import SE3  # noqa: E402
from util import geodesic_rad, rel_translation_err  # noqa: E402

from util import ROT_TOL_RAD, TRANS_REL_TOL, callables, compare, golden_names, load_golden  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vican_b200 import _cabi
    return _cabi.lib()


def _arrays(g, filt=True):
    keep = g.reproj < 0.5 if filt else np.ones(g.n_edges, bool)
    cam, time, marker = g.cam[keep], g.time[keep], g.marker[keep]
    uc, ci = np.unique(cam, return_inverse=True)
    ut, ti = np.unique(time, return_inverse=True)
    return dict(cam=ci.astype(np.int32), time=ti.astype(np.int32), marker=marker.astype(np.int32), R=g.R[keep],
                t=g.t[keep], k_r=g.w[keep], k_t=2.0 * g.w[keep], n_c=len(uc), n_t=len(ut))


def _device_graph(g, a, **kw):
    from vican_b200.solver import DeviceGraph
    C_m = np.transpose(g.marker_R, (0, 2, 1)) @ g.marker_R[0]
    return DeviceGraph(a["cam"], a["time"], a["marker"], a["R"], a["k_r"], a["k_t"], C_m, a["n_c"], a["n_t"], **kw)


def _oracle_pairs(g, a):
    blk = orc.fold_blocks(a["R"], a["k_r"], a["marker"], g.marker_R, 0)
    return orc.aggregate_pairs(a["cam"].astype(np.int64), a["time"].astype(np.int64), blk, a["k_r"], a["n_t"])


# ------------------------------------------------------------------------------- kernels
def test_svd3_factors_kernel_vs_numpy(cuda):
    from vican_b200 import ops
    rng = np.random.default_rng(0)
    M = rng.standard_normal((4096, 3, 3)) * rng.uniform(0.1, 50.0, (4096, 1, 1))
    M[:64, :, 0] *= -1
    rot, sp, si = [x.cpu().numpy() for x in ops.svd3_factors_batch(M)]
    r0, U, S = orc.svd_polar_batch(M)
    Ut = np.transpose(U, (0, 2, 1))
    p0 = (U * S[:, None, :]) @ Ut
    i0 = (U * (1 / S)[:, None, :]) @ Ut
    cond = (S[:, 0] / S[:, 2])[:, None, None]
    assert np.all(np.abs(rot - r0) < 1e-13 * cond)
    assert np.all(np.abs(sp - p0) < 1e-13 * cond * S[:, :1, None])
    assert np.all(np.abs(si - i0) < 1e-13 * cond * cond / S[:, 2:, None])
    pol = ops.polar_so3_batch(M).cpu().numpy()
    assert np.abs(pol - rot).max() == 0.0


def test_se3_batch_kernels_vs_container(cuda):
    from vican_b200 import ops
    rng = np.random.default_rng(1)
    n = 500
    Ra, Rb = syn.random_rotations(rng, n), syn.random_rotations(rng, n)
    ta, tb = rng.normal(0, 3, (n, 3)), rng.normal(0, 3, (n, 3))
    Ri, ti = ops.se3_invert_batch(Ra, ta, round_f32=True)
    Ri, ti = Ri.cpu().numpy(), ti.cpu().numpy()
    Rc, tc = ops.se3_compose_batch(Ra, ta, Rb, tb)
    Rc, tc = Rc.cpu().numpy(), tc.cpu().numpy()
    for i in range(0, n, 7):
        a, b = SE3(R=Ra[i], t=ta[i]), SE3(R=Rb[i], t=tb[i])
        inv = a.inv()                                    # float32 store, as the reference
        assert np.array_equal(Ri[i].astype(np.float32), inv.R())
        assert np.abs(ti[i].astype(np.float32) - inv.t()).max() <= 1e-6 * np.abs(inv.t()).max()
        assert np.abs(Rc[i] - Ra[i] @ Rb[i]).max() < 1e-15 * 4
        assert np.abs(tc[i] - (Ra[i] @ tb[i] + ta[i])).max() < 1e-14
    Ri64, ti64 = ops.se3_invert_batch(Ra, ta)
    assert np.abs(Ri64.cpu().numpy() - np.transpose(Ra, (0, 2, 1))).max() == 0.0
    assert np.abs(ti64.cpu().numpy() + np.einsum("nji,nj->ni", Ra, ta)).max() < 1e-14


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("shape", [(12, 80, 4, 4, 2), (20, 300, 6, 7, 3), (40, 500, 24, 20, 10), (7, 50, 3, 7, 3)])
def test_ingestion_matches_oracle_aggregation(cuda, shape, shuffle):
    g = syn.make_camera_network(3, *shape, outlier_frac=0.1)
    a = _arrays(g)
    if shuffle:   # detections in arbitrary order: exercises the radix-sort path (sorted input skips it)
        perm = np.random.default_rng(0).permutation(a["cam"].shape[0])
        for k in ("cam", "time", "marker", "R", "t", "k_r", "k_t"):
            a[k] = np.ascontiguousarray(a[k][perm])
    dg = _device_graph(g, a)
    pc, pt, B, av = _oracle_pairs(g, a)
    assert dg.n_edges == pc.shape[0]
    # oracle pairs are in first-occurrence order; device pairs are sorted by (time, cam)
    order = np.lexsort((pc, pt))
    assert np.array_equal(dg.t_cam.cpu().numpy(), pc[order])
    assert np.array_equal(dg.t_time.cpu().numpy(), pt[order])
    assert np.abs(dg.t_B.cpu().numpy().reshape(-1, 3, 3) - B[order]).max() < 1e-13
    assert np.abs(dg.t_a.cpu().numpy() - av[order]).max() < 1e-13
    rp = dg.t_rowptr.cpu().numpy()
    assert rp[0] == 0 and rp[-1] == dg.n_edges and np.all(np.diff(rp) == np.bincount(pt, minlength=a["n_t"]))
    cord = dg.c_order.cpu().numpy()                       # camera-pass order: (time window, camera, time)
    assert sorted(cord.tolist()) == list(range(dg.n_edges))
    assert np.array_equal(dg.c_time.cpu().numpy(), pt[order][cord])
    # the camera-pass copy stores the blocks transposed (conflict-free column reads in the pass)
    assert np.array_equal(dg.c_B.cpu().numpy().reshape(-1, 3, 3), np.transpose(dg.t_B.cpu().numpy()[cord].reshape(-1, 3, 3), (0, 2, 1)))
    assert np.array_equal(dg.c_w.cpu().numpy(), dg.t_w.cpu().numpy()[cord])
    deg_t = np.zeros(a["n_t"]); np.add.at(deg_t, pt, av)
    deg_c = np.zeros(a["n_c"]); np.add.at(deg_c, pc, av)
    assert np.abs(dg.deg_t.cpu().numpy() - deg_t).max() < 1e-12
    assert np.abs(dg.deg_c.cpu().numpy() - deg_c).max() < 1e-11
    # tiles cover every camera-pass edge exactly once, one camera per tile, contiguous + sentinel
    nt = dg.n_tiles
    tc = dg.tile_cam.cpu().numpy()[:nt]
    tsent = dg.tile_start.cpu().numpy()[:nt + 1]          # contiguous tiles + sentinel
    ts, te = tsent[:-1], tsent[1:]
    assert np.all(te > ts) and np.all(te - ts <= dg.tile_len)
    assert ts[0] == 0 and te[-1] == dg.n_edges
    # tile_off: first tile of every (window, camera) run -- the fixed order of the per-camera combine
    toff = dg.tile_off.cpu().numpy()
    assert toff.shape[0] == dg.n_windows * a["n_c"] + 1 and toff[0] == 0 and toff[-1] == nt and np.all(np.diff(toff) >= 0)
    for seg in range(dg.n_windows * a["n_c"]):
        assert np.all(tc[toff[seg]:toff[seg + 1]] == seg % a["n_c"])
    cam_of_pos = pc[order][cord]
    for k in range(nt):
        assert np.all(cam_of_pos[ts[k]:te[k]] == tc[k])
    # c_segptr: run (w, c) holds exactly camera c's edges of time window w, in time order
    sp = dg.c_segptr.cpu().numpy()
    n_c, n_w = a["n_c"], dg.n_windows
    assert sp.shape[0] == n_w * n_c + 1 and sp[0] == 0 and sp[-1] == dg.n_edges and np.all(np.diff(sp) >= 0)
    time_of_pos = pt[order][cord]
    for seg in range(n_w * n_c):
        lo, hi = sp[seg], sp[seg + 1]
        assert np.all(cam_of_pos[lo:hi] == seg % n_c)
        assert np.all(np.diff(time_of_pos[lo:hi]) > 0)
        assert np.all(time_of_pos[lo:hi] * n_w // a["n_t"] == seg // n_c)


@pytest.mark.parametrize("shape,tile_len", [((12, 80, 4, 4, 2), None), ((30, 400, 6, 11, 2), 24), ((25, 200, 5, 25, 2), 48),
                                            ((90, 300, 3, 83, 1), 120), ((40, 30, 2, 37, 2), None)])
def test_edge_passes_match_numpy(cuda, shape, tile_len):
    from vican_b200.solver import _ptr, _stream
    g = syn.make_camera_network(4, *shape)
    a = _arrays(g)
    dg = _device_graph(g, a, tile_len=tile_len)
    pc, pt, B, av = _oracle_pairs(g, a)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((a["n_c"], 3, 3))
    lamT = rng.standard_normal((a["n_t"], 3, 3))
    Z0 = dm.pass_time(pc, pt, B, X, a["n_t"])
    W0 = lamT @ Z0
    Y0 = dm.pass_cam(pc, pt, B, W0, a["n_c"])
    Xd = torch.as_tensor(X.reshape(-1, 9)).cuda()
    gs = cuda.vb_gather_stride()                       # padded gather layout: 3 rows x 4 doubles (+ padding to a 128-byte line)
    X12 = torch.empty((a["n_c"], gs), dtype=torch.float64, device="cuda")
    assert cuda.vb_pad_blocks(_ptr(Xd), _ptr(X12), a["n_c"], _stream()) == 0
    assert np.array_equal(X12.cpu().numpy()[:, :12].reshape(-1, 3, 4)[:, :, :3], X)
    Ld = torch.as_tensor(lamT.reshape(-1, 9)).cuda()
    out = torch.zeros((a["n_t"], gs), dtype=torch.float64, device="cuda")
    unpad = lambda t: t.cpu().numpy()[:, :12].reshape(-1, 3, 4)[:, :, :3]  # noqa: E731
    assert cuda.vb_pass_time(C.byref(dg.cgraph), 1, _ptr(X12), None, _ptr(out), _stream()) == 0
    assert np.abs(unpad(out) - Z0).max() < 1e-12 * np.abs(Z0).max()
    assert cuda.vb_pass_time(C.byref(dg.cgraph), 0, _ptr(X12), _ptr(Ld), _ptr(out), _stream()) == 0
    assert np.abs(unpad(out) - W0).max() < 1e-12 * np.abs(W0).max()
    Y = torch.zeros((a["n_c"], 9), dtype=torch.float64, device="cuda")
    assert cuda.vb_pass_cam(C.byref(dg.cgraph), _ptr(out), _ptr(Y), _stream()) == 0
    assert np.abs(Y.cpu().numpy().reshape(-1, 3, 3) - Y0).max() < 1e-12 * np.abs(Y0).max()


# ------------------------------------------------------------------------ rotation stage
@pytest.mark.parametrize("seed,shape,maxiter,outl,filt", [
    (5, (12, 80, 4, 4, 2), 1, 0.0, True), (5, (12, 80, 4, 4, 2), 2, 0.0, True), (5, (12, 80, 4, 4, 2), 3, 0.0, True),
    (6, (12, 80, 4, 4, 2), 10, 0.0, True), (1, (20, 300, 6, 7, 3), 4, 0.0, True),
    (8, (15, 120, 6, 5, 3), 5, 0.2, True), (9, (15, 120, 6, 5, 3), 5, 0.1, False),
    (11, (200, 1500, 24, 20, 10), 6, 0.0, True), (4, (600, 6000, 1, 40, 1), 4, 0.0, True),
    (2, (3, 30, 2, 3, 2), 3, 0.0, True)])
def test_rotation_stage_matches_oracle(cuda, seed, shape, maxiter, outl, filt):
    from vican_b200.solver import solve_rotations
    g = syn.make_camera_network(seed, *shape, outlier_frac=outl, cube=(shape[2] == 24))
    a = _arrays(g, filt)
    dg = _device_graph(g, a)
    pc, pt, B, av = _oracle_pairs(g, a)
    r_c0, r_t0 = orc.so3sync(pc, pt, B, av, a["n_c"], a["n_t"], maxiter)
    rot = solve_rotations(dg, maxiter)
    assert rot.status == 0, rot.status
    r_c = rot.r_c.cpu().numpy().reshape(-1, 3, 3)
    r_t = rot.r_t.cpu().numpy().reshape(-1, 3, 3)
    ec, et = geodesic_rad(r_c, r_c0).max(), geodesic_rad(r_t, r_t0).max()
    assert ec <= ROT_TOL_RAD and et <= ROT_TOL_RAD, (ec, et, list(rot.stats.inner_per_outer[:maxiter]))
    # much tighter in practice: the eigen-solve is converged to 1e-11 relative
    assert ec <= 1e-8 and et <= 1e-8, (ec, et)
    assert rot.stats.time_passes > 0 and rot.stats.cam_passes > 0


# ------------------------------------------------------------------------------ full API
@pytest.mark.parametrize("name", golden_names())
def test_api_matches_reference_golden(cuda, name):
    from vican_b200.bipgo import bipartite_se3sync, object_bipartite_se3sync
    g, params, filter_on, ref = load_golden(name)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(filter_on)
    if g.kind == "object":
        out = object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **params)
    else:
        out = bipartite_se3sync(edges, constraints, nr, nt, ef, dtype=np.float64, **params)
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD, rot
    assert tr <= TRANS_REL_TOL, tr


@pytest.mark.parametrize("cfg,scale,solver", [("cfg1", 0.2, "direct"), ("cfg1", 0.2, "conjugate_gradient"),
                                               ("cfg2", 0.25, "conjugate_gradient"), ("cfg2", 0.1, "direct"),
                                               ("cfg3", 0.03, "conjugate_gradient")])
def test_api_matches_oracle_on_baseline_shapes(cuda, cfg, scale, solver):
    """BASELINE.json config shapes (scaled so the oracle finishes in seconds)."""
    from vican_b200 import bipgo
    g, p = syn.make_config(cfg, scale)
    p["lsqr_solver"] = solver
    p["maxiter"] = min(p["maxiter"], 6)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    if g.kind == "object":
        out = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **p)
        ref, info = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, se3_cls=SE3, return_info=True, **p)
    else:
        out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, dtype=np.float64, **p)
        ref, info = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, return_info=True, **p)
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD and tr <= TRANS_REL_TOL, (rot, tr, bipgo.last_info)
    if solver == "direct":
        assert bipgo.last_info["trans_iters"] == info["itn"] and bipgo.last_info["trans_istop"] == info["istop"]


def test_cg_replays_scipy_iteration_count(cuda):
    """The device CG must stop at the same iteration as scipy's cg (rtol=1e-5 before each step)."""
    import scipy.sparse.linalg as spl
    from vican_b200 import bipgo
    g = syn.make_camera_network(21, 25, 400, 6, 6, 3)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    count = {"n": 0}
    orig = orc._scipy_cg

    def counting_cg(A, b, **kw):
        def cb(xk):
            count["n"] += 1
        return spl.cg(A, b, callback=cb, **kw)
    orc._scipy_cg = counting_cg
    try:
        ref = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, 4, "conjugate_gradient")
    finally:
        orc._scipy_cg = orig
    out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, 4, "conjugate_gradient", dtype=np.float64)
    assert bipgo.last_info["trans_iters"] == count["n"], (bipgo.last_info["trans_iters"], count["n"])
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD and tr <= TRANS_REL_TOL


def test_output_contract(cuda):
    """Keys, container type, dtypes and gauge as the reference returns them."""
    from vican_b200.bipgo import bipartite_se3sync, large_bipartite_so3sync
    g = syn.make_camera_network(13, 12, 40, 3, 4, 2)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    out = bipartite_se3sync(edges, constraints, nr, nt, ef, 3, "conjugate_gradient")
    cams = {k[0] for k in edges}
    times = {k[1].split("_")[0] + "_0" for k in edges}
    assert set(out.keys()) == cams | times
    assert list(out.keys()) == sorted(out.keys())
    v = out[sorted(cams)[0]]
    assert isinstance(v, SE3) and v.R().dtype == np.float32 and v.t().dtype == np.float64
    assert v.inv().R().dtype == np.float32 and (v @ v.inv()).R().shape == (3, 3)
    rots = large_bipartite_so3sync(edges, constraints, nr, ef, 3, dtype=np.float64)
    assert set(rots.keys()) == set(out.keys())
    # gauge: the first camera (lexicographic) is the identity up to the primal/dual updates
    with pytest.raises(ValueError):
        bipartite_se3sync(edges, constraints, nr, nt, ef, 3, "cholesky")
    bad = dict(edges)
    k0 = next(iter(bad))
    bad[(k0[0], k0[1].split("_")[0] + "_99")] = bad[k0]
    with pytest.raises(KeyError):
        bipartite_se3sync(bad, constraints, nr, nt, ef, 3, "conjugate_gradient")


def test_accurate_mode_is_close_to_exact_minimiser(cuda):
    """mode='accurate' (Jacobi-PCG to 1e-12) vs a dense least-squares solve of the same system."""
    from vican_b200 import bipgo
    g = syn.make_camera_network(17, 10, 60, 3, 4, 2)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, 4, "conjugate_gradient", dtype=np.float64,
                                  mode="accurate")
    # exact min-norm solution from the oracle's J, t~ (same rotations to 1e-11)
    ref, info = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, 4, "conjugate_gradient", return_info=True)
    keys = sorted(ref.keys())
    t_acc = np.stack([out[k].t() for k in keys])
    t_ref = np.stack([ref[k][1] for k in keys])
    # reference CG is only 1e-5-accurate; the accurate mode differs from it by the truncation error
    assert rel_translation_err(t_acc, t_ref).max() < 5e-2
    # and has zero mean (minimum-norm gauge of the singular normal equations)
    assert np.abs(t_acc.mean(axis=0)).max() < 1e-8 * np.abs(t_acc).max()


@pytest.mark.parametrize("shape", [(10, 60, 3, 4, 2), (75, 400, 4, 9, 2), (300, 900, 2, 12, 1)])
def test_dense_direct_path_is_the_exact_minimiser(cuda, shape):
    """lsqr_solver='direct', mode='accurate': closed-form elimination of the time nodes + blocked
    Cholesky of the camera Schur complement (csrc/schur.cuh, no cuSOLVER) against a dense
    minimum-norm least-squares solve of the SAME system (oracle J, t~ built from the device's
    rotations).  300 cameras = 10 Cholesky tiles with a ragged last one."""
    from vican_b200 import solver
    g = syn.make_camera_network(23, *shape)
    a = _arrays(g)
    C_m = np.transpose(g.marker_R, (0, 2, 1)) @ g.marker_R[0]
    con = {str(m): SE3(R=g.marker_R[m], t=g.marker_t[m]) for m in range(g.n_markers)}
    t_inv0 = np.stack([np.asarray((con[str(m)].inv() @ con["0"]).t(), np.float64) for m in range(g.n_markers)])
    r_0m = np.transpose(g.marker_R[0])[None] @ g.marker_R
    marker_q = np.einsum("mij,mj->mi", r_0m, t_inv0)
    res = {}
    for solver_name, mode in (("direct", "accurate"), ("conjugate_gradient", "accurate")):
        r = solver.solve_arrays(a["cam"], a["time"], a["marker"], a["R"], a["t"], a["k_r"], a["k_t"], C_m, marker_q,
                                a["n_c"], a["n_t"], 3, solver_name, mode=mode)
        res[solver_name] = r
    r = res["direct"]
    n_c, n_t = a["n_c"], a["n_t"]
    J, tt = orc.translation_system(a["cam"].astype(np.int64), a["time"].astype(np.int64), a["marker"].astype(np.int64),
                                   a["t"], a["k_t"], g.marker_R, t_inv0, 0, r.Rw_c.cpu().numpy(), r.Rw_t.cpu().numpy(),
                                   n_c, n_t, np.arange(n_c), n_c + np.arange(n_t))
    # min-norm minimiser through the scalar Laplacian: J^T J = L (x) I_3
    A = (J.T @ J).toarray()[::3, ::3]
    b = (J.T @ tt).reshape(-1, 3)
    x = np.linalg.pinv(A, rcond=1e-12, hermitian=True) @ b
    x_dev = np.concatenate([r.x_c.cpu().numpy(), r.x_t.cpu().numpy()])
    assert rel_translation_err(x_dev, x).max() < 1e-9
    assert np.abs(x_dev.mean(axis=0)).max() < 1e-10 * np.abs(x_dev).max()
    # stationarity of the normal equations and agreement with the iterative accurate mode
    assert np.abs(A @ x_dev - b).max() < 1e-9 * np.abs(b).max()
    x_pcg = np.concatenate([res["conjugate_gradient"].x_c.cpu().numpy(), res["conjugate_gradient"].x_t.cpu().numpy()])
    assert rel_translation_err(x_pcg, x_dev).max() < 1e-7


def test_dense_direct_path_reports_disconnected_graphs(cuda):
    from vican_b200 import solver
    from vican_b200._cabi import VbError
    g = syn.make_camera_network(3, 8, 40, 2, 3, 2)
    a = _arrays(g)
    # two components: cameras 0..3 only see even time nodes, 4..7 odd ones
    keep = (a["cam"] < 4) == (a["time"] % 2 == 0)
    for k in ("cam", "time", "marker", "R", "t", "k_r", "k_t"):
        a[k] = a[k][keep]
    C_m = np.transpose(g.marker_R, (0, 2, 1)) @ g.marker_R[0]
    G = solver.DeviceGraph(a["cam"], a["time"], a["marker"], a["R"], a["k_r"], a["k_t"], C_m, a["n_c"], a["n_t"])
    rot = solver.RotationResult(torch.eye(3, dtype=torch.float64, device="cuda").reshape(1, 9).repeat(a["n_c"], 1),
                                torch.eye(3, dtype=torch.float64, device="cuda").reshape(1, 9).repeat(a["n_t"], 1), None, 0)
    with pytest.raises(VbError) as ei:
        solver.solve_translations(G, rot, a["t"], np.zeros((g.n_markers, 3)), "direct", mode="accurate")
    assert ei.value.code == 4


def test_float32_call_of_the_notebook(cuda):
    """main.ipynb cell 7 calls ``bipartite_se3sync(..., dtype=np.float32)``.  The reference then runs its sparse
    algebra and ARPACK in single precision and returns float32 rotations with float64 translations; its result is
    1.1e-7 rad / 2.3e-6 (relative translation) away from its own float64 result on these inputs (measured,
    tests/golden/make_golden_f32.py).  Here the arithmetic stays fp64 and ``dtype`` selects the output dtype:
    same dtypes, rotations within the 1e-6 rad contract of the float32 reference, translations within the float32
    reference's own distance to float64."""
    import os
    from vican_b200 import bipgo
    from util import GOLDEN_DIR
    g, params, filter_on, ref64 = load_golden("net_small_cg_it3")
    z = np.load(os.path.join(GOLDEN_DIR, "f32_net_small_cg_it3.npz"))
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(filter_on)
    out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, dtype=np.float32, **params)
    keys = [str(k) for k in z["out_keys"]]
    assert sorted(str(k) for k in out.keys()) == keys
    k0 = keys[0]
    assert out[k0].R().dtype == np.dtype(str(z["R_dtype"])) == np.float32
    assert out[k0].t().dtype == np.dtype(str(z["t_dtype"])) == np.float64
    Ra = np.stack([np.asarray(out[k].R(), np.float64) for k in keys])
    ta = np.stack([out[k].t() for k in keys])
    assert geodesic_rad(Ra, z["out_R"]).max() <= ROT_TOL_RAD
    assert rel_translation_err(ta, z["out_t"]).max() <= 1e-5
    # and it is the float64 result rounded: closer to the float64 reference than the float32 reference is
    R64 = np.stack([ref64[k][0] for k in keys])
    assert geodesic_rad(Ra, R64).max() <= 1.2e-7
    assert rel_translation_err(ta, np.stack([ref64[k][1] for k in keys])).max() <= TRANS_REL_TOL


def test_verbose_reports_the_references_eigenvalue_readout(cuda, capsys):
    """bipgo.py:288-292, :336-339: the five eigenvalues nearest zero of every outer iteration (three ~0 and
    lambda_4, lambda_5) and eigengap = |lambda_4 / lambda_3|, against the oracle's dense eigen-solve."""
    from vican_b200 import bipgo
    g = syn.make_camera_network(5, 15, 120, 4, 5, 2)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, 5, "conjugate_gradient", dtype=np.float64, verbose=True)
    printed = capsys.readouterr().out
    assert printed.count("eigengap=") == 5 and "evals0=" in printed
    ref, info = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, 5, "conjugate_gradient", return_info=True)
    rot, tr = compare(out, ref)
    assert rot <= 1e-8 and tr <= TRANS_REL_TOL            # the diagnostics do not disturb the solve
    ev = np.asarray(bipgo.last_info["evals"])
    assert ev.shape == (5, 5) and bipgo.last_info["n_components"] == 1 and not bipgo.last_info["early_exit"]
    for it in range(5):
        o = np.sort(np.abs(info["evals"][it]))            # oracle: 5 nearest sigma (3 tiny, then lambda_4, lambda_5)
        d = np.sort(np.abs(ev[it]))
        scale = o[4]
        assert np.all(np.abs(d[:3] - o[:3]) <= 1e-9 * scale), (it, d, o)
        assert np.all(np.abs(d[3:] - o[3:]) <= 1e-6 * scale), (it, d, o)


def test_early_exit_on_a_disconnected_graph_like_the_reference(cuda):
    """bipgo.py:283-284: the loop stops once max |lambda_1..5| <= 1e-6, which needs lambda_4, lambda_5 ~ 0, i.e. a
    graph with more than one component (six zero modes).  Same stopping iteration as the oracle."""
    import warnings
    from vican_b200 import bipgo
    g = syn.make_camera_network(1, 12, 80, 3, 8, 2, sigma_R=0.0, sigma_t=0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    # two components: cameras 0..5 only see timesteps 0..39, cameras 6..11 timesteps 40..79 (noise-free, so
    # all six zero modes are exact and the reference leaves its loop after the first iteration)
    edges = {k: v for k, v in edges.items() if (int(k[0]) < 6) == (int(k[1].split("_")[0]) < 40)}
    nr, nt, ef = callables(True)
    maxiter = 12
    ref, info = orc.bipartite_se3sync_oracle(edges, cons, nr, nt, ef, maxiter, "conjugate_gradient", return_info=True)
    n_ref = len(info["evals"])
    assert n_ref < maxiter                                  # the reference's early exit fired
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        out = bipgo.bipartite_se3sync(edges, cons, nr, nt, ef, maxiter, "conjugate_gradient", dtype=np.float64)
    assert any("connected components" in str(w.message) for w in wlist)
    li = bipgo.last_info
    assert li["n_components"] == 2 and li["early_exit"] and li["outer_done"] == n_ref, (li["outer_done"], n_ref, li["evals"])
    assert all(np.all(np.isfinite(v.R())) and np.all(np.isfinite(v.t())) for v in out.values())
    # the gauge camera's component is determined: its cameras agree with the oracle
    for c in range(6):
        assert geodesic_rad(out[str(c)].R(), ref[str(c)][0]).max() <= ROT_TOL_RAD


def test_stalled_eigen_iteration_is_reported(cuda, monkeypatch):
    """An eigen-iteration that stops at its step cap must not pass silently (the reference's ARPACK call raises
    ArpackNoConvergence there): EigenConvergenceWarning by default, ConvergenceError with strict=True."""
    import warnings
    from vican_b200 import bipgo, solver
    g = syn.make_camera_network(9, 12, 60, 3, 4, 2)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    real = solver.solve_rotations
    monkeypatch.setattr(solver, "solve_rotations", lambda gg, maxiter, **kw: real(gg, maxiter, **dict(kw, max_inner=2)))
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, 3, "conjugate_gradient", dtype=np.float64)
    assert any(issubclass(w.category, bipgo.EigenConvergenceWarning) for w in wlist)
    assert bipgo.last_info["eig_status"] == 2 and len(out) == 12 + 60
    with pytest.raises(solver.ConvergenceError):
        bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, 3, "conjugate_gradient", dtype=np.float64, strict=True)


def test_inexact_early_eigen_solves_are_verified_or_repeated(cuda):
    """vb_so3_options.tol_early (profiles/r2_inexact_inner.md): with the default margin the inexact run lands on the
    all-tight result to rounding and on the oracle's; with a margin too short for the tight end-game to find the fixed
    point the run is flagged (stats.inexact_unverified) and repeated all-tight: identical bits."""
    from vican_b200.solver import solve_rotations
    g = syn.make_camera_network(11, 40, 500, 24, 10, 4, outlier_frac=0.1, cube=True)
    a = _arrays(g, False)                                   # outliers left in: slow outer convergence
    dg = _device_graph(g, a)
    maxiter = 10
    tight = solve_rotations(dg, maxiter, tol_early=0.0)
    dflt = solve_rotations(dg, maxiter)                     # tol_early = 1e-5, the last 4 iterations tight
    assert not dflt.repeated_tight and dflt.stats.inexact_unverified == 0
    assert sum(dflt.stats.inner_per_outer[:maxiter]) < sum(tight.stats.inner_per_outer[:maxiter])
    assert geodesic_rad(dflt.r_c.cpu().numpy(), tight.r_c.cpu().numpy()).max() <= 1e-12
    assert geodesic_rad(dflt.r_t.cpu().numpy(), tight.r_t.cpu().numpy()).max() <= 1e-12
    pc, pt, B, av = _oracle_pairs(g, a)
    r_c0, r_t0 = orc.so3sync(pc, pt, B, av, a["n_c"], a["n_t"], maxiter)
    assert geodesic_rad(dflt.r_c.cpu().numpy().reshape(-1, 3, 3), r_c0).max() <= 1e-8
    assert geodesic_rad(dflt.r_t.cpu().numpy().reshape(-1, 3, 3), r_t0).max() <= 1e-8
    short = solve_rotations(dg, maxiter, tol_early=1e-2, early_margin=1)
    assert short.repeated_tight
    assert torch.equal(short.r_c, tight.r_c) and torch.equal(short.r_t, tight.r_t)
