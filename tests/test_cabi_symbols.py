"""CPU: the C-ABI shared library builds, loads and exports every symbol include/vican_b200.h
declares (no compute calls -- there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "vican_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", src))
    names -= {"vb_allreduce_fn"}
    return sorted(names)


@pytest.fixture(scope="module")
def lib():
    from vican_b200 import _cabi
    if not os.path.exists(_cabi.SO_PATH):
        _cabi.build()
    return _cabi.load_library()


def test_header_declares_expected_surface():
    syms = header_symbols()
    for must in ("vb_so3sync_run", "vb_pass_time", "vb_pass_cam", "vb_ingest_build", "vb_trans_cg", "vb_trans_lsqr",
                 "vb_se3_compose_batch", "vb_se3_invert_batch", "vb_polar_so3_batch", "vb_nccl_allreduce"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from vican_b200 import _cabi
    for name in header_symbols():
        assert hasattr(lib, name), name
        assert name in _cabi.SIGNATURES, "ctypes signature missing for " + name
    for name in _cabi.SIGNATURES:
        assert name in header_symbols(), "bound symbol not declared in the header: " + name


def test_version_and_status_strings(lib):
    assert b"sm_100a" in lib.vb_version()
    assert lib.vb_status_string(0) == b"ok"
    assert b"converge" in lib.vb_status_string(1)


def test_struct_layouts_match_header():
    from vican_b200._cabi import VbGraph, VbSo3Options, VbSo3Stats
    assert ctypes.sizeof(VbGraph) == 5 * 8 + 21 * 8
    assert ctypes.sizeof(VbSo3Options) == 4 + 4 + 8 + 8 + 8 + 8 + 8 + 8 + 8 + 8
    assert ctypes.sizeof(VbSo3Stats) == 6 * 4 + 3 * 8 + 3 * 8 + 8 + 64 * 4 + 8 + 8 + 4 + 4 + 4 + 4 + 64 * 5 * 8 + 8


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vican_b200 import _cabi
    with pytest.raises(RuntimeError):
        _cabi.lib()
