"""CPU checks of the DEVICE algorithm as stated in numpy (oracle/device_model.py): matrix-free two-pass
L-apply + width-3 LOBPCG, one-hop spanning start, and the primal multiply without edge passes once the
eigen-iteration accepts its start block -- against the reference-faithful oracle (dense eigh)."""
import numpy as np
import pytest

from oracle import device_model as dm
from oracle import vican_oracle as orc
from vican_b200 import synthetic as syn
from util import geodesic_rad


def _pairs(seed, shape, outl=0.0):
    g = syn.make_camera_network(seed, *shape, outlier_frac=outl)
    keep = g.reproj < 0.5
    uc, ci = np.unique(g.cam[keep], return_inverse=True)
    ut, ti = np.unique(g.time[keep], return_inverse=True)
    blk = orc.fold_blocks(g.R[keep], g.w[keep], g.marker[keep], g.marker_R, 0)
    pc, pt, B, a = orc.aggregate_pairs(ci, ti, blk, g.w[keep], len(ut))
    return pc, pt, B, a, len(uc), len(ut)


@pytest.mark.parametrize("seed,shape,outl,maxiter", [(5, (12, 80, 4, 4, 2), 0.0, 8), (8, (15, 120, 6, 5, 3), 0.2, 12)])
def test_device_algorithm_matches_oracle_with_and_without_the_pass_savers(seed, shape, outl, maxiter):
    pc, pt, B, a, n_c, n_t = _pairs(seed, shape, outl)
    r_c0, r_t0 = orc.so3sync(pc, pt, B, a, n_c, n_t, maxiter)
    res = {}
    for short, span in ((False, False), (True, True)):
        st = dm.LobpcgStats()
        r_c, r_t = dm.so3sync_model(pc, pt, B, a, n_c, n_t, maxiter, tol=1e-13, stats=st, shortcut=short,
                                    spanning_start=span)
        res[short] = (r_c, r_t, st)
        assert geodesic_rad(r_c, r_c0).max() < 1e-9 and geodesic_rad(r_t, r_t0).max() < 1e-9
    # same iterates to rounding, fewer passes
    assert geodesic_rad(res[True][0], res[False][0]).max() < 1e-11
    p_on, p_off = res[True][2].passes, res[False][2].passes
    assert sum(t for t, _ in p_on) < sum(t for t, _ in p_off)
    # converged outer iterations: 1 camera pass (eigen-residual) + 1 time pass (dual gather), nothing else
    assert p_on[-1] == (1, 1) and p_off[-1] == (2, 2)
    # the spanning start never needs more eigen-steps than identity blocks (here: strictly fewer passes in outer 0
    # even after paying its own two)
    assert p_on[0][0] <= p_off[0][0] + 1


def test_shortcut_identity_holds_for_any_invertible_mix():
    """The algebra behind the shortcut: for R with SO(3) blocks and any invertible C,
    project_SO3((R C)_c (R C)_0^-1) = R_c R_0^T, and P Lambda_T P^T (R R_0^T) = (P Lambda_T P^T R) R_0^T."""
    pc, pt, B, a, n_c, n_t = _pairs(3, (9, 50, 3, 4, 2))
    rng = np.random.default_rng(0)
    R = syn.random_rotations(rng, n_c)
    C = rng.standard_normal((3, 3)) + 2 * np.eye(3)
    V = R @ C
    Z = V @ np.linalg.inv(V[0])
    rot, _, _ = dm.svd_factors(Z)
    assert np.abs(rot - R @ R[0].T).max() < 1e-13
    LamT = rng.uniform(0.5, 2.0, n_t)[:, None, None] * np.eye(3)[None]
    ppwr = lambda X: dm.pass_cam(pc, pt, B, LamT @ dm.pass_time(pc, pt, B, X, n_t), n_c)  # noqa: E731
    Y = ppwr(R)
    assert np.abs(ppwr(rot) - Y @ R[0].T).max() < 1e-12 * np.abs(Y).max()


@pytest.mark.parametrize("seed,shape,outl,maxiter", [(5, (12, 80, 4, 4, 2), 0.0, 10), (8, (15, 120, 6, 5, 3), 0.2, 12),
                                                      (3, (20, 150, 4, 8, 2), 0.1, 10)])
def test_inexact_early_eigen_solves_reach_the_same_fixed_point(seed, shape, outl, maxiter):
    """vb_so3_options.tol_early: the eigen-solves of the early outer iterations stop at 1e-5 instead of 1e-13; the
    last four are tight.  The primal-dual iteration contracts so strongly that the result equals the reference-
    faithful oracle's (exact eigen-solves throughout) to 1e-9 rad, with far fewer applications of L -- and the
    criterion the device uses to trust such a run (the last two tight iterations accept their start block at the
    first step) holds."""
    pc, pt, B, a, n_c, n_t = _pairs(seed, shape, outl)
    r_c0, r_t0 = orc.so3sync(pc, pt, B, a, n_c, n_t, maxiter)
    st_t, st_e = dm.LobpcgStats(), dm.LobpcgStats()
    dm.so3sync_model(pc, pt, B, a, n_c, n_t, maxiter, tol=1e-13, stats=st_t)
    r_c, r_t = dm.so3sync_model(pc, pt, B, a, n_c, n_t, maxiter, tol=1e-13, stats=st_e, tol_early=1e-5, early_margin=4)
    assert geodesic_rad(r_c, r_c0).max() < 1e-9 and geodesic_rad(r_t, r_t0).max() < 1e-9
    assert st_e.applies[-1] == 1 and st_e.applies[-2] == 1
    assert sum(st_e.applies) < sum(st_t.applies)
