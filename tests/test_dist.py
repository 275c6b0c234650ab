"""Multi-rank coverage: world_size-2 gloo test of the host-side sharding logic on CPU, and (on a
box with >= 2 GPUs) edge-sharded solve == single-GPU solve through NCCL."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WORKER = os.path.join(HERE, "dist_worker.py")


def _run(mode, nproc, port, collective=None):
    env = dict(os.environ)
    env.setdefault("TQDM_DISABLE", "1")
    if collective is not None:
        env["VICAN_B200_COLLECTIVE"] = collective
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)


def test_shard_range_partitions():
    from vican_b200.dist import shard_range
    for n in (1, 7, 100, 1_000_003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_sharded_passes_match_unsharded():
    out = _run("cpu", 2, 29533)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_CPU_OK" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("collective", ["peer", "nccl"])
def test_two_gpu_edge_sharded_solve_matches_single_gpu(collective):
    """peer: camera pass fused with the one-shot all-reduce over NVLink peer windows (+ the
    standalone one-shot kernel against NCCL); nccl: separate ncclAllReduce per camera pass."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run("gpu", 2, 29534 if collective == "peer" else 29536, collective)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_GPU_OK" in out.stdout, out.stdout[-2000:]
    if collective == "peer":
        assert "PEER_ALLREDUCE_OK" in out.stdout
