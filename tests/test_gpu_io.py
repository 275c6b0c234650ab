"""GPU tests of the data formats either side of the solver, through the C ABI and against
the oracle: ``EdgeTable.from_arrays``, ``io.load_edges`` (the notebook's ``cam_marker_edges.pt``),
``io.EdgeAccumulator`` and the device-side incremental ingestion (``solver.StreamingGraph`` / ``io.DeviceStream``)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import vican_oracle as orc           # noqa: E402
from vican_b200 import io as vio                 # noqa: E402
from vican_b200 import synthetic as syn          # noqa: E402
from vican_b200.geometry import SE3              # noqa: E402

from util import ROT_TOL_RAD, TRANS_REL_TOL, callables, compare, geodesic_rad, rel_translation_err  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vican_b200 import _cabi
    return _cabi.lib()


@pytest.fixture(scope="module")
def case():
    g = syn.make_camera_network(31, 14, 160, 5, 6, 3, outlier_frac=0.1)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    ref = orc.bipartite_se3sync_oracle(edges, cons, nr, nt, ef, 5, "conjugate_gradient")
    return g, edges, cons, (nr, nt, ef), ref


def test_solve_from_preevaluated_arrays(cuda, case):
    """EdgeTable.from_arrays: ids as strings, poses / weights as arrays, no Python callable runs."""
    from vican_b200 import bipgo
    g, edges, cons, (nr, nt, ef), ref = case
    kept = [(k, v) for k, v in edges.items() if ef(v)]
    tab = bipgo.EdgeTable.from_arrays([k[0] for k, _ in kept], [k[1] for k, _ in kept],
                                      np.stack([v["pose"].R() for _, v in kept]), np.stack([v["pose"].t() for _, v in kept]),
                                      np.array([nr(v) for _, v in kept]), np.array([nt(v) for _, v in kept]), cons)
    out = bipgo.solve_table(tab, 5, "conjugate_gradient", dtype=np.float64)
    rot, tr = compare(out, ref)
    assert rot <= 1e-8 and tr <= TRANS_REL_TOL, (rot, tr)


def test_solve_from_a_notebook_style_pt_file(cuda, case, tmp_path):
    """main.ipynb cells 3/5: torch.save(edges) -> io.load_edges -> bipartite_se3sync."""
    from vican_b200 import bipgo
    g, edges, cons, (nr, nt, ef), ref = case
    path = str(tmp_path / "cam_marker_edges.pt")
    vio.save_edges(edges, path)
    loaded = vio.load_edges(path)
    assert list(loaded.keys()) == list(edges.keys())
    out = bipgo.bipartite_se3sync(loaded, cons, nr, nt, ef, 5, "conjugate_gradient", dtype=np.float64)
    rot, tr = compare(out, ref)
    assert rot <= 1e-8 and tr <= TRANS_REL_TOL, (rot, tr)


def test_solve_from_the_accumulator(cuda, case):
    """io.EdgeAccumulator fed image by image (cam.py:101-184), then one solve from its arrays."""
    from vican_b200 import bipgo
    g, edges, cons, (nr, nt, ef), ref = case
    acc = vio.EdgeAccumulator(nr, nt, ef)
    chunks = {}
    for k, v in edges.items():
        chunks.setdefault(v["im_filename"], {})[k] = v
    for c in chunks.values():
        acc.add(c)
    out = bipgo.solve_table(acc.table(cons), 5, "conjugate_gradient", dtype=np.float64)
    rot, tr = compare(out, ref)
    assert rot <= 1e-8 and tr <= TRANS_REL_TOL, (rot, tr)


def test_streaming_graph_appends_without_resorting(cuda):
    """solver.StreamingGraph: the graph built by three appends of time-node ranges holds the same time-sorted
    arrays as the one-shot ingestion, its camera-pass windows are the chunks' windows, and both solve to the
    same poses (<= 1e-10 rad; the camera-side sums run over a different window partition)."""
    from vican_b200 import solver
    g = syn.make_camera_network(8, 30, 900, 4, 9, 2)
    a = dict(cam=g.cam.astype(np.int32), time=g.time.astype(np.int32), marker=g.marker.astype(np.int32), R=g.R.reshape(-1, 9),
             t=g.t, k_r=g.w, k_t=2.0 * g.w)
    C_m = np.transpose(g.marker_R, (0, 2, 1)) @ g.marker_R[0]
    con = {str(m): SE3(R=g.marker_R[m], t=g.marker_t[m]) for m in range(g.n_markers)}
    t_inv0 = np.stack([np.asarray((con[str(m)].inv() @ con["0"]).t(), np.float64) for m in range(g.n_markers)])
    marker_q = np.einsum("mij,mj->mi", np.transpose(g.marker_R[0])[None] @ g.marker_R, t_inv0)
    n_c, n_t = 30, 900
    one = solver.DeviceGraph(a["cam"], a["time"], a["marker"], a["R"], a["k_r"], a["k_t"], C_m, n_c, n_t)
    sg = solver.StreamingGraph(n_c, C_m)
    for lo, hi in ((0, 250), (250, 640), (640, 900)):
        m = (a["time"] >= lo) & (a["time"] < hi)
        sg.append(a["cam"][m], a["time"][m] - lo, a["marker"][m], a["R"][m], a["t"][m], a["k_r"][m], a["k_t"][m], hi - lo)
    gs = sg.graph()
    assert (gs.n_edges, gs.n_t, gs.n_raw) == (one.n_edges, one.n_t, one.n_raw)
    for name in ("t_rowptr", "t_cam", "t_time", "t_B", "t_a", "t_w", "pair_start", "deg_t"):
        assert torch.equal(getattr(gs, name), getattr(one, name)), name
    assert (gs.deg_c - one.deg_c).abs().max() <= 1e-12 * one.deg_c.abs().max()
    assert torch.equal(torch.sort(gs.c_order).values, torch.arange(gs.n_edges, device="cuda", dtype=torch.int32))
    assert torch.equal(gs.c_time, gs.t_time[gs.c_order.long()])
    r1, r2 = solver.solve_rotations(one, 5), solver.solve_rotations(gs, 5)
    assert geodesic_rad(r1.r_c.cpu().numpy(), r2.r_c.cpu().numpy()).max() <= 1e-10
    assert geodesic_rad(r1.r_t.cpu().numpy(), r2.r_t.cpu().numpy()).max() <= 1e-10
    t1 = solver.solve_translations(one, r1, a["t"], marker_q, "conjugate_gradient")
    t2 = solver.solve_translations(gs, r2, sg.t, marker_q, "conjugate_gradient")
    assert t1.iters == t2.iters
    assert rel_translation_err(t2.x_c.cpu().numpy(), t1.x_c.cpu().numpy()).max() <= 1e-7
    assert rel_translation_err(t2.x_t.cpu().numpy(), t1.x_t.cpu().numpy()).max() <= 1e-7
    for solver_name in ("direct",):                         # LSQR rows index the appended raw arrays
        u1 = solver.solve_translations(one, r1, a["t"], marker_q, solver_name)
        u2 = solver.solve_translations(gs, r2, sg.t, marker_q, solver_name)
        assert u1.iters == u2.iters and rel_translation_err(u2.x_t.cpu().numpy(), u1.x_t.cpu().numpy()).max() <= 1e-6


def test_device_stream_image_by_image_matches_the_oracle(cuda):
    """io.DeviceStream: per-image dictionaries arriving in time order are appended to the device graph in
    chunks of complete timesteps; the solve after the last image equals the oracle's on the whole dictionary."""
    g = syn.make_camera_network(12, 11, 150, 4, 5, 2)
    edges, cons = syn.to_edge_dict(g, SE3)
    nr, nt, ef = callables(True)
    ref = orc.bipartite_se3sync_oracle(edges, cons, nr, nt, ef, 4, "conjugate_gradient")
    ds = vio.DeviceStream(sorted({k[0] for k in edges}), cons, nr, nt, ef, chunk_detections=300)
    images = {}
    for k, v in edges.items():                              # synthetic dictionaries are time-major already
        images.setdefault(v["im_filename"], {})[k] = v
    n_appends = 0
    for chunk in images.values():
        before = ds.graph.n_t
        ds.add(chunk)
        n_appends += ds.graph.n_t != before
    out = ds.solve(4, "conjugate_gradient", dtype=np.float64)
    assert n_appends >= 3 and ds.n_late == 0
    rot, tr = compare(out, ref)
    assert rot <= 1e-8 and tr <= TRANS_REL_TOL, (rot, tr)
