"""Pin the CPU oracle against outputs of the REAL reference (tests/golden/*.npz,
made by tests/golden/make_golden.py in the build container)."""
import numpy as np
import pytest

from oracle import vican_oracle as orc
from vican_b200 import synthetic as syn
from vican_b200.geometry import SE3

from util import ROT_TOL_RAD, TRANS_REL_TOL, callables, compare, golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    g, params, filter_on, ref = load_golden(name)
    edges, constraints = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(filter_on)
    if g.kind == "object":
        out = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, se3_cls=SE3, **params)
    else:
        out = orc.bipartite_se3sync_oracle(edges, constraints, nr, nt, ef, **params)
    rot, tr = compare(out, ref)
    # rotations: the restatement reproduces the reference to rounding (dense eigh vs ARPACK)
    assert rot < 1e-9, rot
    # translations: cg replays to ~1e-11; scipy's lsqr amplifies 1e-16 input perturbations to
    # ~1e-7 on the object-calibration graphs (measured: identical J, b differing by 3e-14 ->
    # x differing by 3e-8 with the same itn), so only the stated parity tolerance is asserted
    if params["lsqr_solver"] == "conjugate_gradient":
        assert tr < 1e-9, tr
    assert rot < ROT_TOL_RAD and tr < TRANS_REL_TOL, (rot, tr)


def test_golden_set_not_empty():
    assert len(golden_names()) >= 8
