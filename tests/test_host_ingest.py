"""Host side of the ingestion: dictionary flatten, pre-evaluated array
path, streaming accumulator, reader of the notebook's ``cam_marker_edges.pt``.  CPU only."""
import sys
import types

import numpy as np
import pytest

from vican_b200 import io as vio
from vican_b200 import synthetic as syn
from vican_b200.bipgo import EdgeTable
from vican_b200.geometry import SE3

FIELDS = ("cam_ids", "time_ids", "marker_ids", "cam_idx", "time_idx", "marker_idx", "R", "t", "k_r", "k_t",
          "markerC", "marker_q", "round_kr_f32", "root", "n_raw")


def _graph(outliers=0.2):
    return syn.make_camera_network(3, 12, 40, 4, 5, 2, outlier_frac=outliers)


def _same(a, b):
    for f in FIELDS:
        x, y = np.asarray(getattr(a, f)), np.asarray(getattr(b, f))
        assert x.dtype == y.dtype and np.array_equal(x, y), f


def test_node_order_is_the_references_lexicographic_order():
    """bipgo.py:225-229: nodes = np.unique(strings) -> '10' sorts before '2'.  Time nodes follow the
    order of the translation unknowns, np.unique over t+'_0' (bipgo.py:420-430: '10_0' < '1_0')."""
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    tab = EdgeTable(edges, cons, nr, nt, ef)
    assert list(tab.cam_ids) == sorted({k[0] for k in edges})
    assert list(tab.time_ids) == sorted({k[1].split("_")[0] for k in edges}, key=lambda t: t + "_0")
    assert list(tab.time_ids).index("10") < list(tab.time_ids).index("1")
    # unknown_index() = position in np.unique over all node names together
    unk_c, unk_t = tab.unknown_index()
    names = np.unique(np.asarray(list(tab.cam_ids) + [t + "_0" for t in tab.time_ids]))
    assert [names[i] for i in unk_c] == list(tab.cam_ids)
    assert [names[i] for i in unk_t] == [t + "_0" for t in tab.time_ids]
    assert np.all(np.diff(unk_c) > 0) and np.all(np.diff(unk_t) > 0)
    assert list(tab.cam_ids).index("10") < list(tab.cam_ids).index("2")
    # every detection is coded with the index of its own ids
    keys = list(edges.keys())
    for e in (0, len(keys) // 2, len(keys) - 1):
        assert tab.cam_ids[tab.cam_idx[e]] == keys[e][0]
        assert tab.time_ids[tab.time_idx[e]] == keys[e][1].split("_")[0]
        assert tab.marker_ids[tab.marker_idx[e]] == keys[e][1].split("_")[1]
        assert np.array_equal(tab.R[e].reshape(3, 3), edges[keys[e]]["pose"].R())


def test_callables_see_only_kept_detections_and_once():
    g = _graph(0.3)
    edges, cons = syn.to_edge_dict(g, SE3)
    calls = {"f": 0, "r": 0, "t": 0}

    def ef(e):
        calls["f"] += 1
        return e["reprojected_err"] < 0.5

    def nr(e):
        assert e["reprojected_err"] < 0.5
        calls["r"] += 1
        return e["w"]

    def nt(e):
        assert e["reprojected_err"] < 0.5
        calls["t"] += 1
        return 2.0 * e["w"]

    tab = EdgeTable(edges, cons, nr, nt, ef)
    assert calls["f"] == len(edges) and calls["r"] == tab.n_raw and calls["t"] == tab.n_raw
    assert 0 < tab.n_raw < len(edges)


def test_missing_constraint_is_a_keyerror_and_empty_is_a_valueerror():
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    bad = dict(cons)
    del bad[sorted(bad)[-1]]
    with pytest.raises(KeyError):
        EdgeTable(edges, bad, nr, nt, ef)
    with pytest.raises(ValueError):
        EdgeTable(edges, cons, nr, nt, lambda e: False)


def test_from_arrays_equals_dict_flatten():
    g = _graph(0.2)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    tab = EdgeTable(edges, cons, nr, nt, ef)
    kept = [(k, v) for k, v in edges.items() if ef(v)]
    tab2 = EdgeTable.from_arrays([k[0] for k, _ in kept], [k[1] for k, _ in kept],
                                 np.stack([v["pose"].R() for _, v in kept]), np.stack([v["pose"].t() for _, v in kept]),
                                 np.array([nr(v) for _, v in kept]), np.array([nt(v) for _, v in kept]), cons)
    for f in FIELDS:
        if f == "round_kr_f32":
            continue
        x, y = np.asarray(getattr(tab, f)), np.asarray(getattr(tab2, f))
        assert np.array_equal(x, y), f


def test_accumulator_streams_per_image_chunks():
    g = _graph(0.2)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    tab = EdgeTable(edges, cons, nr, nt, ef)
    acc = vio.EdgeAccumulator(nr, nt, ef)
    # chunks = detections of one image (one camera at one timestep), as estimate_pose_worker returns them
    chunks = {}
    for k, v in edges.items():
        chunks.setdefault(v["im_filename"], {})[k] = v
    assert acc.add(None) == 0                       # images without detections yield None (cam.py:139)
    n = sum(acc.add(c) for c in chunks.values())
    assert n == tab.n_raw == len(acc)
    tab2 = acc.table(cons)
    # same multiset of detections; the order follows arrival, so compare through a canonical sort
    def canon(t):
        key = np.lexsort((t.marker_idx, t.cam_idx, t.time_idx))
        return t.cam_idx[key], t.time_idx[key], t.marker_idx[key], t.R[key], t.t[key], t.k_r[key], t.k_t[key]
    for x, y in zip(canon(tab), canon(tab2)):
        assert np.array_equal(x, y)
    # a re-detected key replaces the older detection instead of duplicating it
    k0 = next(k for k, v in edges.items() if ef(v))
    assert acc.add({k0: edges[k0]}) == 1 and len(acc) == tab.n_raw


def test_marker_whitelist_like_estimate_pose_mp():
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    acc = vio.EdgeAccumulator(nr, nt, ef, marker_ids=["0", "1"])
    acc.add(edges)
    assert len(acc) == sum(1 for k in edges if k[1].split("_")[1] in ("0", "1"))


def test_load_edges_reads_reference_pickles_without_the_reference_package(tmp_path):
    """main.ipynb cell 3/5: torch.save(dict with vican.geometry.SE3 values).  The loader maps that
    class onto ours, so /root/reference need not be importable."""
    import torch

    # a stand-in with the reference container's attribute layout (geometry.py:194-218)
    pkg, mod = types.ModuleType("vican"), types.ModuleType("vican.geometry")

    class RefSE3(object):
        def __init__(self, R, t):
            self._R, self._t = R, t
            self._pose = np.eye(4, dtype=np.float32)
            self._pose[:3, :3] = R
            self._pose[:3, 3] = t

    RefSE3.__name__ = RefSE3.__qualname__ = "SE3"
    RefSE3.__module__ = "vican.geometry"
    mod.SE3 = RefSE3
    pkg.geometry = mod
    sys.modules["vican"], sys.modules["vican.geometry"] = pkg, mod
    try:
        g = _graph(0.0)
        src = {}
        for e in range(20):
            src[(str(g.cam[e]), "%d_%d" % (g.time[e], g.marker[e]))] = {
                "pose": RefSE3(g.R[e].copy(), g.t[e].copy()), "corners": np.zeros((4, 2)),
                "reprojected_err": 0.01, "im_filename": "x"}
        path = str(tmp_path / "cam_marker_edges.pt")
        torch.save(src, path)
    finally:
        del sys.modules["vican"], sys.modules["vican.geometry"]
    out = vio.load_edges(path)
    assert list(out.keys()) == list(src.keys())
    for k in src:
        assert type(out[k]["pose"]) is SE3
        assert np.array_equal(out[k]["pose"].R(), src[k]["pose"]._R)
        assert np.array_equal(out[k]["pose"].t(), src[k]["pose"]._t)
        assert np.allclose(out[k]["pose"].inv().R(), src[k]["pose"]._R.T, atol=1e-6)


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/vican"), reason="reference checkout not present (GPU box)")
def test_load_edges_reads_a_pickle_written_with_the_real_reference_class(tmp_path):
    """Same as above with the REAL ``vican.geometry.SE3`` (build container only): write with the reference
    importable, read back in a subprocess-free way after hiding it again."""
    import torch
    sys.path.insert(0, "/root/reference")
    try:
        from vican.geometry import SE3 as RefSE3
        g = _graph(0.0)
        src = {(str(g.cam[e]), "%d_%d" % (g.time[e], g.marker[e])): {
            "pose": RefSE3(R=g.R[e].copy(), t=g.t[e].copy()), "corners": np.zeros((4, 2)), "reprojected_err": 0.01,
            "im_filename": "x"} for e in range(15)}
        src[("0", "999_0")] = {"pose": RefSE3(pose=np.eye(4)), "corners": np.zeros((4, 2)), "reprojected_err": 0.0,
                               "im_filename": "y"}
        path = str(tmp_path / "cam_marker_edges.pt")
        torch.save(src, path)
        want = {k: (np.array(v["pose"].R()), np.array(v["pose"].t()), np.array(v["pose"].inv().R())) for k, v in src.items()}
    finally:
        sys.path.remove("/root/reference")
        for m in [m for m in sys.modules if m == "vican" or m.startswith("vican.")]:
            del sys.modules[m]
    out = vio.load_edges(path)
    assert list(out) == list(want)
    for k, (R, t, Ri) in want.items():
        assert type(out[k]["pose"]) is SE3
        assert np.array_equal(out[k]["pose"].R(), R) and np.array_equal(out[k]["pose"].t(), t)
        assert np.array_equal(out[k]["pose"].inv().R(), Ri)      # same float32 rounding as the reference container


def test_accumulator_retires_a_detection_whose_newer_version_is_filtered_out():
    """cam.py:263 merges per-image dictionaries: a re-detected key REPLACES the older detection; if the newer one
    fails edge_filter the solver must not keep using the older one (the reference would filter the merged entry)."""
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    acc = vio.EdgeAccumulator(nr, nt, ef)
    acc.add(edges)
    n0 = len(acc)
    k0 = next(iter(edges))
    bad = dict(edges[k0], reprojected_err=1.0)           # newer detection of the same key, rejected by the filter
    assert acc.add({k0: bad}) == 0 and len(acc) == n0 - 1
    merged = dict(edges)
    merged[k0] = bad
    tab_ref = EdgeTable(merged, cons, nr, nt, ef)         # what the reference would solve on
    tab = acc.table(cons)
    assert tab.n_raw == tab_ref.n_raw == n0 - 1
    # a later accepted detection of that key revives it
    assert acc.add({k0: edges[k0]}) == 1 and len(acc) == n0


def test_mixed_pose_dtypes_are_rejected():
    """The float32 rounding of `k_r * pose.R()` (geometry.py:209-211) is mirrored per call, so a dictionary that
    mixes float32 and float64 pose arrays is refused instead of silently promoted."""
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    k0 = next(iter(edges))
    p = edges[k0]["pose"]
    edges[k0] = dict(edges[k0], pose=SE3(R=p.R().astype(np.float32), t=p.t()))
    with pytest.raises(ValueError):
        EdgeTable(edges, cons, nr, nt, ef)


# ---------------------------------------------------------------------------------------------
# the C walk of the dictionary (vican_b200/csrc/flatten.c) against the list-comprehension statement
# ---------------------------------------------------------------------------------------------

def _both(edges, cons, nr, nt, ef, monkeypatch):
    from vican_b200 import bipgo
    assert bipgo._vb_flatten is not None, "vican_b200/_vb_flatten*.so not built (run __graft_entry__.build())"
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "0")
    a = EdgeTable(edges, cons, nr, nt, ef)
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "1")
    b = EdgeTable(edges, cons, nr, nt, ef)
    return a, b


def test_c_flatten_equals_python_flatten(monkeypatch):
    g = _graph(0.3)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    _same(*_both(edges, cons, nr, nt, ef, monkeypatch))
    # weights of other numeric types: Python int, numpy scalars, 0-d arrays
    _same(*_both(edges, cons, lambda e: 2, lambda e: np.float32(0.5), lambda e: np.bool_(True), monkeypatch))
    _same(*_both(edges, cons, lambda e: np.float64(1.5), lambda e: np.asarray(0.25), lambda e: 1, monkeypatch))


def test_c_flatten_float32_poses_and_column_translations(monkeypatch):
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    e32 = {k: dict(v, pose=SE3(R=v["pose"].R().astype(np.float32), t=v["pose"].t().astype(np.float32).reshape(3, 1)))
           for k, v in edges.items()}
    a, b = _both(e32, cons, nr, nt, ef, monkeypatch)
    _same(a, b)
    assert a.round_kr_f32 and a.R.dtype == np.float64
    # numpy-scalar weights keep the product in float64 (bipgo.py:212 under numpy's promotion rules)
    a, b = _both(e32, cons, lambda e: np.float64(1.0), nt, ef, monkeypatch)
    assert not a.round_kr_f32 and not b.round_kr_f32
    # non-contiguous views and other dtypes go through numpy's conversion
    big = np.zeros((6, 6))
    odd = {}
    for k, v in edges.items():
        big = np.zeros((6, 6))
        big[::2, ::2] = v["pose"].R()
        odd[k] = dict(v, pose=SE3(R=big[::2, ::2], t=v["pose"].t()))
    a = EdgeTable(odd, cons, nr, nt, ef)
    c = EdgeTable(edges, cons, nr, nt, ef)
    _same(a, c)


def test_c_flatten_errors_match(monkeypatch):
    g = _graph(0.0)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    keys = list(edges.keys())
    mixed = dict(edges)
    mixed[keys[3]] = dict(edges[keys[3]], pose=SE3(R=edges[keys[3]]["pose"].R().astype(np.float32),
                                                   t=edges[keys[3]]["pose"].t()))
    for flag in ("0", "1"):
        monkeypatch.setenv("VICAN_B200_PY_FLATTEN", flag)
        with pytest.raises(ValueError, match="mix float32 and float64"):
            EdgeTable(mixed, cons, nr, nt, ef)
        with pytest.raises(ValueError, match="no edge passes"):
            EdgeTable(edges, cons, nr, nt, lambda e: False)
        with pytest.raises(ZeroDivisionError):
            EdgeTable(edges, cons, lambda e: 1 / 0, nt, ef)
        with pytest.raises(KeyError):
            EdgeTable(edges, {k: v for k, v in cons.items() if k != "1"}, nr, nt, ef)
        with pytest.raises(KeyError):                                    # detection without a pose
            EdgeTable({keys[0]: {"corners": edges[keys[0]]["corners"], "reprojected_err": 0.0}}, cons, nr, nt, ef)


def test_c_flatten_call_counts_and_order(monkeypatch):
    g = _graph(0.3)
    edges, cons = syn.to_edge_dict(g, SE3)
    _, _, ef0 = syn.default_callables()
    seen = {"f": [], "r": [], "t": []}
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "0")
    EdgeTable(edges, cons, lambda e: seen["r"].append(id(e)) or 1.0, lambda e: seen["t"].append(id(e)) or 1.0,
              lambda e: seen["f"].append(id(e)) or ef0(e))
    kept = [id(v) for v in edges.values() if ef0(v)]
    assert seen["f"] == [id(v) for v in edges.values()]                  # insertion order, once each
    assert seen["r"] == kept and seen["t"] == kept


def test_object_rekey_in_c_equals_the_python_loop(monkeypatch):
    """object_bipartite_se3sync's re-key + invert step (bipgo.py:526-531): pose stacking and the new dictionary are built
    in C (csrc/flatten.c: poses, rekey); same dictionary as the Python loop -- same key order, same field OBJECTS,
    float32 pose containers with the same numbers.  The device batch inversion is replaced by numpy here (CPU test)."""
    import torch
    from vican_b200 import bipgo, ops

    def fake_invert(R, t, round_f32=False):
        R = np.asarray(R, np.float64).reshape(-1, 3, 3)
        t = np.asarray(t, np.float64).reshape(-1, 3)
        Ri = np.ascontiguousarray(np.transpose(R, (0, 2, 1)))
        ti = -(Ri @ t[:, :, None])[:, :, 0]
        if round_f32:
            Ri, ti = Ri.astype(np.float32).astype(np.float64), ti.astype(np.float32).astype(np.float64)
        return torch.from_numpy(Ri), torch.from_numpy(ti)

    monkeypatch.setattr(ops, "se3_invert_batch", fake_invert)
    g = syn.make_object_calibration(seed=3, n_times=60, n_markers=10, min_visible=3, max_visible=10)
    edges, _ = syn.to_edge_dict(g, SE3)
    keys = list(edges.keys())
    edges[keys[5]] = dict(edges[keys[5]], pose=SE3(R=edges[keys[5]]["pose"].R().astype(np.float32),
                                                   t=edges[keys[5]]["pose"].t()))          # np.stack would upcast it
    root = str(min(int(k[1].split("_")[1]) for k in edges))
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "0")
    assert bipgo._vb_flatten is not None
    a = bipgo._rekey_inverted(edges, root)
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "1")
    b = bipgo._rekey_inverted(edges, root)
    assert list(a.keys()) == list(b.keys()) and len(a) == len(edges)
    for (ka, va), (kb, vb), v0 in zip(a.items(), b.items(), edges.values()):
        assert ka == kb and isinstance(ka, tuple) and list(va.keys()) == list(vb.keys()) == ["pose", "corners", "reprojected_err", "im_filename"]
        assert va["corners"] is v0["corners"] and va["im_filename"] is v0["im_filename"] and va["reprojected_err"] is v0["reprojected_err"]
        for f in ("R", "t"):
            x, y = getattr(va["pose"], f)(), getattr(vb["pose"], f)()
            assert x.dtype == y.dtype == np.float32 and np.array_equal(x, y)
        assert va["pose"]._pose.dtype == np.float32 and np.array_equal(va["pose"]._pose, vb["pose"]._pose)
        assert np.array_equal((va["pose"] @ va["pose"].inv())._pose, (vb["pose"] @ vb["pose"].inv())._pose)
    # the full flatten of both dictionaries agrees as well
    nr, nt, ef = syn.default_callables()
    cons = {root: SE3(pose=np.eye(4))}
    _same(EdgeTable(a, cons, nr, nt, ef), EdgeTable(b, cons, nr, nt, ef))
    # errors of the Python loop: a missing field, a key that is not "t_m"
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "0")
    bad = dict(edges)
    bad[keys[0]] = {k: v for k, v in edges[keys[0]].items() if k != "corners"}
    with pytest.raises(KeyError):
        bipgo._rekey_inverted(bad, root)
    with pytest.raises(ValueError):
        bipgo._rekey_inverted({(keys[0][0], "7"): edges[keys[0]]}, root)
    # float32 poses take the container's own inv() (numpy float32 arithmetic), C re-key only
    e32 = {k: dict(v, pose=SE3(R=v["pose"].R().astype(np.float32), t=v["pose"].t().astype(np.float32))) for k, v in edges.items()}
    c = bipgo._rekey_inverted(e32, root)
    monkeypatch.setenv("VICAN_B200_PY_FLATTEN", "1")
    d = bipgo._rekey_inverted(e32, root)
    assert list(c.keys()) == list(d.keys())
    assert all(np.array_equal(x["pose"]._pose, y["pose"]._pose) for x, y in zip(c.values(), d.values()))


def test_object_host_path_end_to_end_against_the_oracle(monkeypatch):
    """The host side of object_bipartite_se3sync (re-key + invert + hand-over to bipartite_se3sync + result filter,
    bipgo.py:524-543) with the two device calls replaced -- batch inversion by numpy, the solve by the oracle's
    bipartite_se3sync -- must reproduce the oracle's object_bipartite_se3sync on the original dictionary."""
    import torch
    from oracle import vican_oracle as orc
    from vican_b200 import bipgo, ops
    from util import compare

    def fake_invert(R, t, round_f32=False):
        R = np.asarray(R, np.float64).reshape(-1, 3, 3)
        t = np.asarray(t, np.float64).reshape(-1, 3)
        Ri = np.ascontiguousarray(np.transpose(R, (0, 2, 1)))
        ti = -(Ri @ t[:, :, None])[:, :, 0]
        if round_f32:
            Ri, ti = Ri.astype(np.float32).astype(np.float64), ti.astype(np.float32).astype(np.float64)
        return torch.from_numpy(Ri), torch.from_numpy(ti)

    def oracle_solve(edges, constraints, noise_model_r, noise_model_t, edge_filter, maxiter, lsqr_solver, dtype=np.float32, **kw):
        out = orc.bipartite_se3sync_oracle(edges, constraints, noise_model_r, noise_model_t, edge_filter, maxiter, lsqr_solver)
        return {k: SE3(R=R.astype(dtype), t=t) for k, (R, t) in out.items()}

    monkeypatch.setattr(ops, "se3_invert_batch", fake_invert)
    monkeypatch.setattr(bipgo, "bipartite_se3sync", oracle_solve)
    g = syn.make_object_calibration(seed=2, n_times=60, n_markers=10, min_visible=3, max_visible=10)
    edges, _ = syn.to_edge_dict(g, SE3)
    nr, nt, ef = syn.default_callables()
    ref = orc.object_bipartite_se3sync_oracle(edges, nr, nt, ef, 3, "conjugate_gradient", se3_cls=SE3)
    for flag in ("0", "1"):
        monkeypatch.setenv("VICAN_B200_PY_FLATTEN", flag)
        out = bipgo.object_bipartite_se3sync(edges, nr, nt, ef, maxiter=3, lsqr_solver="conjugate_gradient", dtype=np.float64)
        assert sorted(out.keys()) == sorted(ref.keys()) and all("_" not in k for k in out)
        rot, tr = compare(out, ref)
        assert rot <= 1e-12 and tr <= 1e-10, (flag, rot, tr)
