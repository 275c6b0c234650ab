"""Parity at the FULL size of the BASELINE.json configurations (run on the B200 box, ``-m gpu``).

cfg1 / cfg2 / cfg3 run through the dictionary drop-in against the oracle on the same dictionaries;
cfg5 (maxiter = 500) at 10 % of its time nodes; cfg4 at 1/20 scale (500 cameras, 50 000 time nodes,
2.5 M edges -- the largest graph the oracle's dense eigen-solves finish in minutes) through the array
API against ``solve_arrays_oracle``.  Tolerances are BASELINE.json's (1e-6 rad, 1e-6
relative translation per node); cfg2 -- the graph on which the reference's truncated CG is chaotic
in the rounding of its mat-vec -- is held to 2e-7, which needs scipy's row-sum ORDER (csrc/cg.cuh).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import vican_oracle as orc           # noqa: E402
from vican_b200 import synthetic as syn          # noqa: E402
from vican_b200.geometry import SE3              # noqa: E402

from util import (ROT_TOL_RAD, TRANS_REL_TOL, callables, compare, geodesic_rad, load_full_golden,  # noqa: E402
                  rel_translation_err)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vican_b200 import _cabi
    return _cabi.lib()


def _run_dict(cfg, scale, solver=None, maxiter=None):
    from vican_b200 import bipgo
    g, p = syn.make_config(cfg, scale)
    if solver is not None:
        p["lsqr_solver"] = solver
    if maxiter is not None:
        p["maxiter"] = maxiter
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    if g.kind == "object":
        out = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **p)
        info_dev = dict(bipgo.last_info)
        ref, info = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, se3_cls=SE3, return_info=True, **p)
    else:
        out = bipgo.bipartite_se3sync(edges, constraints, nr, nt, ef, dtype=np.float64, **p)
        info_dev = dict(bipgo.last_info)
        ref, info = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, return_info=True, **p)
    return out, ref, info_dev, info, (edges, constraints, p, g)


def _vs_real_reference(out, golden, g, tr_tol, oracle_out=None, rot_tol=1e-9):
    """The same result against the answer of the REAL reference on the same (regenerated, digest-checked) inputs
    (tests/golden/make_golden_fullsize.py); with ``oracle_out`` the oracle is pinned to it as well."""
    ref, _ = load_full_golden(golden, g)
    rot, tr = compare(out, ref)
    assert rot <= rot_tol and tr <= tr_tol, (golden, rot, tr)
    if oracle_out is not None:
        rot, tr = compare(oracle_out, ref)
        assert rot <= rot_tol and tr <= tr_tol, ("oracle", golden, rot, tr)


@pytest.mark.parametrize("solver", ["direct", "conjugate_gradient"])
def test_cfg1_full_size(cuda, solver):
    """small_room shape: 20 cameras, 5 000 timesteps, 105 000 detections, maxiter 10."""
    out, ref, info_dev, info, ctx = _run_dict("cfg1", 1.0, solver)
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD and tr <= TRANS_REL_TOL, (rot, tr, info_dev)
    assert rot <= 1e-9 and tr <= 2e-7, (rot, tr)
    if solver == "direct":
        assert info_dev["trans_iters"] == info["itn"] and info_dev["trans_istop"] == info["istop"]
    # (scipy's lsqr amplifies rounding-level differences of its inputs to ~1e-7: only the contract is asserted for it)
    _vs_real_reference(out, "full_cfg1_direct" if solver == "direct" else "full_cfg1_cg", ctx[3],
                       TRANS_REL_TOL if solver == "direct" else 2e-7)


def test_cfg2_full_size_and_bitwise_reproducible(cuda):
    """cube_calib shape: 24 markers x 2 000 frames, object_bipartite_se3sync, cg, maxiter 4.  scipy's own
    truncated iterate moves by 1e-6 .. 5e-6 when the 24 marker rows of its mat-vec are summed in another
    order (measured); the device replays the order and lands within 2e-7.  The whole solve
    is free of atomics on one GPU: two runs agree bit for bit."""
    from vican_b200 import bipgo
    out, ref, info_dev, info, (edges, _, p, g) = _run_dict("cfg2", 1.0)
    rot, tr = compare(out, ref)
    assert rot <= 1e-9, rot
    assert tr <= 2e-7, (tr, info_dev)
    nr, nt, ef = callables(True)
    out2 = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, dtype=np.float64, **p)
    assert info_dev["trans_iters"] == bipgo.last_info["trans_iters"]
    for k in out:
        assert np.array_equal(out[k].t(), out2[k].t()) and np.array_equal(out[k].R(), out2[k].R()), k
    _vs_real_reference(out, "full_cfg2", g, 2e-7)


def test_cfg3_full_size(cuda):
    """large_shop shape: 200 cameras, 10 000 timesteps, 24-marker cube, 2 M detections, cg, maxiter 10."""
    out, ref, info_dev, info, ctx = _run_dict("cfg3", 1.0)
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD and tr <= TRANS_REL_TOL, (rot, tr, info_dev)
    assert rot <= 1e-9 and tr <= 2e-7, (rot, tr)
    _vs_real_reference(out, "full_cfg3", ctx[3], 2e-7, oracle_out=ref)


def test_cfg5_tenth_convergence_stress(cuda):
    """cfg5: large_shop shape + 20 % outliers removed by edge_filter, maxiter = 500, at 10 % of the time
    nodes (200 cameras, 1 000 timesteps, 200 000 detections): the oracle's 500 dense eigen-solves finish
    in about a minute.  Full API (rotations + cg translations)."""
    out, ref, info_dev, info, ctx = _run_dict("cfg5", 0.1)
    rot, tr = compare(out, ref)
    assert rot <= ROT_TOL_RAD and tr <= TRANS_REL_TOL, (rot, tr, info_dev)
    assert rot <= 1e-8, rot
    # ... and against the REAL reference's 500 iterations on the same inputs
    _vs_real_reference(out, "tenth_cfg5", ctx[3], TRANS_REL_TOL, oracle_out=ref, rot_tol=1e-8)


def test_cfg4_twentieth_scale_matches_oracle(cuda):
    """cfg4 shape at 1/20 scale: 500 cameras, 50 000 single-marker time nodes, 50 cameras per node =
    2.5 M edges, maxiter 10, cg -- solve_arrays vs solve_arrays_oracle on the same arrays."""
    from vican_b200 import solver
    n_c, n_t, d, maxiter = 500, 50_000, 50, 10
    g = syn.make_camera_network(4, n_c, n_t, 1, d, 1)
    assert g.n_edges == n_t * d
    k_r, k_t = g.w, 2.0 * g.w
    C_m = np.transpose(g.marker_R, (0, 2, 1)) @ g.marker_R[0]
    res = solver.solve_arrays(g.cam.astype(np.int32), g.time.astype(np.int32), g.marker.astype(np.int32), g.R, g.t,
                              k_r, k_t, C_m, np.zeros((1, 3)), n_c, n_t, maxiter, "conjugate_gradient")
    ref = orc.solve_arrays_oracle(g.cam, g.time, g.marker, g.R, g.t, k_r, k_t, g.marker_R, np.zeros((1, 3)), 0,
                                  n_c, n_t, maxiter, "conjugate_gradient")
    ec = geodesic_rad(res.Rw_c.cpu().numpy(), ref[0]).max()
    et = geodesic_rad(res.Rw_t.cpu().numpy(), ref[1]).max()
    xc = rel_translation_err(res.x_c.cpu().numpy(), ref[2]).max()
    xt = rel_translation_err(res.x_t.cpu().numpy(), ref[3]).max()
    assert max(ec, et) <= ROT_TOL_RAD and max(xc, xt) <= TRANS_REL_TOL, (ec, et, xc, xt)
    assert max(ec, et) <= 1e-9 and max(xc, xt) <= 2e-7, (ec, et, xc, xt)


def _object_graph_arrays(n_t):
    """Object-calibration shape as a network: markers play the camera role, frames the time role."""
    g = syn.make_object_calibration(0, n_t, 24)
    cam, time = g.marker.astype(np.int32), g.time.astype(np.int32)
    Rinv = np.transpose(g.R, (0, 2, 1)).copy()
    tinv = -np.einsum("eij,ej->ei", Rinv, g.t)
    return cam, time, np.zeros(cam.shape[0], dtype=np.int32), Rinv, tinv, g.w, 2.0 * g.w


def _device_vs_scipy_cg(n_t, perturb=None):
    import scipy.sparse.linalg as spl
    from vican_b200 import solver
    cam, time, zeros, Rinv, tinv, k_r, k_t = _object_graph_arrays(n_t)
    n_c = 24
    dg = solver.DeviceGraph(cam, time, zeros, Rinv, k_r, k_t, np.eye(3).reshape(1, 9), n_c, n_t)
    rot = solver.solve_rotations(dg, 4)
    a = solver.solve_translations(dg, rot, tinv, np.zeros((1, 3)), "conjugate_gradient")
    b = solver.solve_translations(dg, rot, tinv, np.zeros((1, 3)), "conjugate_gradient")
    assert a.iters == b.iters
    assert torch.equal(a.x_c, b.x_c) and torch.equal(a.x_t, b.x_t)      # no atomics: identical bits
    Rw_c, Rw_t = rot.world_rotations()
    J, tt = orc.translation_system(cam.astype(np.int64), time.astype(np.int64), zeros.astype(np.int64), tinv, k_t,
                                   np.eye(3)[None], np.zeros((1, 3)), 0, Rw_c.cpu().numpy(), Rw_t.cpu().numpy(),
                                   n_c, n_t, np.arange(n_c), n_c + np.arange(n_t))
    A, rhs = J.T @ J, J.T @ tt
    count = [0]
    x, code = spl.cg(A, rhs, callback=lambda xk: count.__setitem__(0, count[0] + 1))
    assert code == 0 and a.iters == count[0], (a.iters, count[0])
    x = x.reshape(-1, 3)
    xd = np.concatenate([a.x_c.cpu().numpy(), a.x_t.cpu().numpy()])
    err = rel_translation_err(xd, x).max()
    # scipy's own sensitivity: the same call with the right-hand side moved by one part in 1e16
    rng = np.random.default_rng(0)
    x2, _ = spl.cg(A, rhs * (1.0 + 1e-16 * rng.standard_normal(rhs.shape)))
    own = rel_translation_err(x2.reshape(-1, 3), x).max()
    return err, own


@pytest.mark.parametrize("n_t", [500, 1600, 2000])
def test_cg_replays_scipy_row_order_on_arrays(cuda, n_t):
    """The replayed CSR product through the array API (unknown order: cameras, then time nodes): same
    iteration count as scipy's cg on the explicit J^T J and <= 2e-7 per node on the object-calibration
    shape (any other row-sum order of the 24 marker rows lands ~1e-6 away), identical bits run to run."""
    err, own = _device_vs_scipy_cg(n_t)
    assert err <= 2e-7, (err, own)


def test_cg_on_an_instance_where_scipy_itself_is_chaotic(cuda):
    """24 markers x 1200 frames (seed 0) is an instance on which scipy's truncated iterate moves by
    ~2e-6 when its right-hand side is perturbed by 1e-16 or its dot products are summed in another order
    (measured): no implementation short of scipy's own binary reproduces it to 1e-6.  The device result
    has the same iteration count and stays within a small multiple of scipy's own sensitivity."""
    err, own = _device_vs_scipy_cg(1200)
    assert own > 2e-7                      # the instance really is chaotic (else it belongs in the test above)
    assert err <= 50.0 * own + 2e-7, (err, own)
