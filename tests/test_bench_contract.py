"""bench.py contract pieces that can be checked without a GPU: the reference arm (CPU oracle port) prints exactly
one JSON line on stdout with the keys the driver reads, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(env_extra=None):
    # the contract is checked on a small sample (the driver's run uses cfg4 at 1/10 scale: ~2 min of CPU work)
    env = dict(os.environ, TQDM_DISABLE="1", VICAN_B200_REF_SAMPLE="small", **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--gpus", "1"], capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_json_line():
    out = _run()
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "primal_dual_iter_per_s" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "cfg4" and d["config"]["n_edges"] == 50_000_000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    rs = d["config"]["reference_sample"]
    assert rs["edge_scale_factor"] == 50_000_000 / rs["n_edges"] and cb["sample_iterations"] == rs["maxiter"]
    assert d["e2e"] == {"value": d["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""
