"""Evaluation helpers: ``optimize_gauge_SO3`` / ``optimize_gauge_SE3`` /
``distance_SO3`` / ``angle`` of vican/geometry.py.  CPU: the numpy restatement against goldens
produced by the real reference (tests/golden/make_golden_eval.py).  GPU: the CUDA kernels, through
the C ABI, against those goldens and the restatement."""
import os

import numpy as np
import pytest

from oracle import vican_oracle as orc
from vican_b200 import synthetic as syn
from vican_b200.geometry import SE3

from util import GOLDEN_DIR


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN_DIR, "eval_helpers.npz"))


def test_oracle_gauge_matches_reference(gold):
    gR, gt = orc.optimize_gauge_SE3_oracle(gold["Rgt"], gold["tgt"], gold["Rest"], gold["test"])
    assert np.abs(gR - gold["gauge_R"]).max() < 1e-13
    assert np.abs(gt - gold["gauge_t"]).max() < 1e-12
    gR2, none = orc.optimize_gauge_SE3_oracle(gold["Rgt"], None, gold["Rest"], None)
    assert none is None and np.abs(gR2 - gold["gauge_so3"]).max() < 1e-13


def test_oracle_distance_matches_reference(gold):
    assert np.abs(orc.distance_SO3_oracle(gold["Rgt"], gold["Rest"]) - gold["dist_deg"]).max() < 1e-9
    assert np.abs(orc.angle_deg_oracle(gold["Rest"]) - gold["angle_deg"]).max() < 1e-9
    eye = np.broadcast_to(np.eye(3), gold["Rspecial"].shape)
    assert np.array_equal(orc.distance_SO3_oracle(eye, gold["Rspecial"]), gold["dist_special"])


# ------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vican_b200 import _cabi
    return _cabi.lib()


@pytest.mark.gpu
def test_gauge_kernel_vs_reference_golden(cuda, gold):
    from vican_b200 import ops
    gR, gt = ops.optimize_gauge_batch(gold["Rgt"], gold["tgt"], gold["Rest"], gold["test"])
    assert np.abs(gR.cpu().numpy() - gold["gauge_R"]).max() < 1e-12
    assert np.abs(gt.cpu().numpy() - gold["gauge_t"]).max() < 1e-12
    gR2, none = ops.optimize_gauge_batch(gold["Rgt"], None, gold["Rest"], None)
    assert none is None and np.abs(gR2.cpu().numpy() - gold["gauge_so3"]).max() < 1e-12


@pytest.mark.gpu
def test_distance_kernel_vs_reference_golden(cuda, gold):
    from vican_b200 import ops
    d = ops.distance_so3_batch(gold["Rgt"], gold["Rest"]).cpu().numpy()
    # arccos amplifies the rounding of the trace by 1 / sin(angle): 1e-9 deg covers angles >= 1e-5 deg
    assert np.abs(d - gold["dist_deg"]).max() < 1e-9
    a = ops.distance_so3_batch(gold["Rest"]).cpu().numpy()
    assert np.abs(a - gold["angle_deg"]).max() < 1e-9
    eye = np.ascontiguousarray(np.broadcast_to(np.eye(3), gold["Rspecial"].shape))
    assert np.array_equal(ops.distance_so3_batch(eye, gold["Rspecial"]).cpu().numpy(), gold["dist_special"])


@pytest.mark.gpu
def test_reference_signatures_on_containers(cuda, gold):
    from vican_b200 import geometry as geo
    n = 40
    a = [SE3(R=gold["Rgt"][i], t=gold["tgt"][i]) for i in range(n)]
    b = [SE3(R=gold["Rest"][i], t=gold["test"][i]) for i in range(n)]
    G = geo.optimize_gauge_SE3(a, b)
    gR, gt = orc.optimize_gauge_SE3_oracle(gold["Rgt"][:n], gold["tgt"][:n], gold["Rest"][:n], gold["test"][:n])
    assert np.abs(G.R() - gR).max() < 1e-12 and np.abs(G.t() - gt).max() < 1e-12
    assert np.abs(geo.optimize_gauge_SO3([x.R() for x in a], [x.R() for x in b]) - gR).max() < 1e-12
    assert abs(geo.distance_SO3(gold["Rgt"][3], gold["Rest"][3]) - gold["dist_deg"][3]) < 1e-9
    assert abs(geo.angle(gold["Rest"][5]) - gold["angle_deg"][5]) < 1e-9
    with pytest.raises(AssertionError):
        geo.distance_SO3(np.eye(4), np.eye(3))


@pytest.mark.gpu
def test_large_batch_gauge_is_deterministic_and_recovers_the_gauge(cuda):
    """1 M poses (the parity harness's own inner loop at cfg4 size): exact gauge recovery on
    noise-free data, bitwise reproducible."""
    import torch
    from vican_b200 import ops
    rng = np.random.default_rng(5)
    n = 1_000_000
    Rb = torch.as_tensor(syn.random_rotations(rng, n)).cuda()
    tb = torch.as_tensor(rng.normal(0, 5, (n, 3))).cuda()
    G = syn.random_rotations(rng, 1)[0]
    g = rng.normal(0, 1, 3)
    Gt, gt_ = torch.as_tensor(G).cuda(), torch.as_tensor(g).cuda()
    Ra = Rb @ Gt                       # a = b @ G
    ta = tb + Rb @ gt_
    R1, t1 = ops.optimize_gauge_batch(Ra, ta, Rb, tb)
    R2, t2 = ops.optimize_gauge_batch(Ra, ta, Rb, tb)
    assert torch.equal(R1, R2) and torch.equal(t1, t2)
    assert np.abs(R1.cpu().numpy() - G).max() < 1e-12 and np.abs(t1.cpu().numpy() - g).max() < 1e-10
    d = ops.distance_so3_batch(Ra, Rb @ R1)
    assert float(d.max()) < 1e-5       # degrees; arccos floor of fp64 near the identity (~1e-6 deg)


@pytest.mark.gpu
def test_evaluate_against_mirrors_notebook_cell9(cuda):
    from vican_b200 import geometry as geo
    rng = np.random.default_rng(9)
    n = 30
    Rgt, tgt = syn.random_rotations(rng, n), rng.normal(0, 4, (n, 3))
    Gr, Gt = syn.random_rotations(rng, 1)[0], rng.normal(0, 1, 3)
    # est = Gauge @ gt (a world-frame change), plus noise
    Re = (Gr @ Rgt) @ syn.so3_exp(rng.normal(0, 1e-3, (n, 3)))
    te = (Gr @ tgt.T).T + Gt + rng.normal(0, 1e-3, (n, 3))
    gt = {str(i): SE3(R=Rgt[i], t=tgt[i]) for i in range(n)}
    est = {str(i): SE3(R=Re[i], t=te[i]) for i in range(n)}
    est["extra"] = est["0"]
    keys, r_err, t_err, G = geo.evaluate_against(gt, est)
    assert keys == [str(i) for i in range(n)]
    # the host statement of cell 9 with the containers (float32 4x4 products): agreement to float32 rounding
    Gh_R, Gh_t = orc.optimize_gauge_SE3_oracle(np.transpose(Rgt, (0, 2, 1)), -np.einsum("nji,nj->ni", Rgt, tgt),
                                               np.transpose(Re, (0, 2, 1)), -np.einsum("nji,nj->ni", Re, te))
    assert np.abs(G.R() - Gh_R).max() < 1e-12 and np.abs(G.t() - Gh_t).max() < 1e-11
    assert r_err.max() < 0.5 and t_err.max() < 0.02 and r_err.shape == (n,)
