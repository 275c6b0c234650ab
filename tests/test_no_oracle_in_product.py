"""The oracle is test infrastructure: nothing under vican_b200/ (the product) nor bench.py's GPU arm
may import it, and the product must not carry a CPU fallback."""
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "vican_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                # comments may cite scipy (the algorithm being replaced); no product file may import it
                assert not re.search(r"^\s*(from|import)\s+scipy\b", src, flags=re.M), "scipy imported in " + f


def test_bench_uses_oracle_only_for_cpu_legs():
    """bench.py may run the oracle only as the CPU baseline: in cpu_sample_solve (cpu_baseline / --impl reference
    legs) and in configs_block (the CPU port timed next to the drop-in on the same dictionaries)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import", src)]
    assert len(uses) == 2
    for u in uses:
        fn_start = src.rfind("\ndef ", 0, u) + 1
        assert src[fn_start:].startswith(("def cpu_sample_solve", "def configs_block")), src[fn_start:fn_start + 40]
