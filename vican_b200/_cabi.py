"""ctypes binding of ``libvican_b200.so`` (C ABI declared in ``include/vican_b200.h``).

There is NO CPU fallback: if the extension is missing or no CUDA device is visible,
``lib()`` raises -- the product path must fail loudly rather than compute elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvican_b200.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")
FLATTEN_SRC = os.path.join(_HERE, "csrc", "flatten.c")

_lock = threading.Lock()
_lib = None

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
VP = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F64 = C.c_double

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, VP, VP, I64, VP)


class VbGraph(C.Structure):
    _fields_ = [
        ("n_c", I64), ("n_t", I64), ("n_edges", I64), ("n_tiles", I64), ("n_windows", I64),
        ("t_rowptr", VP), ("t_cam", VP), ("t_B", VP), ("t_w", VP),
        ("c_segptr", VP), ("c_order", VP), ("c_time", VP), ("c_B", VP), ("c_w", VP),
        ("tile_cam", VP), ("tile_start", VP), ("tile_off", VP), ("tile_part", VP),
        ("deg_t", VP), ("deg_c", VP),
        ("st_ptr", VP), ("st_idx", VP), ("st_w", VP), ("sc_ptr", VP), ("sc_idx", VP), ("sc_w", VP),
    ]


class VbArrival(C.Structure):
    _fields_ = [("n_chunks", I32), ("sorted_input", I32), ("h_raw_end", c_i64p), ("h_events", C.POINTER(VP))]


class VbSo3Options(C.Structure):
    _fields_ = [("maxiter", I32), ("max_inner", I32), ("tol", F64), ("allreduce", VP), ("allreduce_ctx", VP),
                ("profile_events", I32), ("no_shortcut", I32), ("identity_start", I32), ("eval_gap", I32),
                ("tol_early", F64), ("early_margin", I32), ("reserved", I32), ("peer_ctx", VP)]


class VbSo3Stats(C.Structure):
    _fields_ = [
        ("outer_done", I32), ("time_passes", I32), ("cam_passes", I32), ("lobpcg_steps", I32),
        ("kernel_launches", I32), ("stalled_outer", I32),
        ("theta", F64 * 3), ("resid", F64 * 3), ("anorm", F64), ("inner_per_outer", I32 * 64),
        ("time_pass_ms", F64), ("cam_pass_ms", F64), ("time_pass_timed", I32), ("cam_pass_timed", I32),
        ("shortcut_outer", I32), ("early_exit", I32), ("evals_hist", (F64 * 5) * 64),
        ("inexact_unverified", I32), ("reserved3", I32),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/vican_b200.h
SIGNATURES = {
    "vb_version": (C.c_char_p, []),
    "vb_launch_count": (I64, []),
    "vb_status_string": (C.c_char_p, [C.c_int]),
    "vb_se3_compose_batch": (C.c_int, [VP, VP, VP, VP, VP, VP, I64, C.c_int, VP]),
    "vb_se3_invert_batch": (C.c_int, [VP, VP, VP, VP, I64, C.c_int, VP]),
    "vb_polar_so3_batch": (C.c_int, [VP, VP, I64, VP]),
    "vb_svd3_factors_batch": (C.c_int, [VP, VP, VP, VP, I64, VP]),
    "vb_gauge_workspace_bytes": (I64, [I64]),
    "vb_optimize_gauge": (C.c_int, [VP, VP, VP, VP, I64, VP, VP, VP, I64, VP]),
    "vb_distance_so3_batch": (C.c_int, [VP, VP, VP, I64, VP]),
    "vb_se3_left_compose_batch": (C.c_int, [VP, VP, VP, VP, VP, VP, I64, C.c_int, VP]),
    "vb_ingest_workspace_bytes": (I64, [I64]),
    "vb_ingest_sort": (C.c_int, [VP, VP, I64, I64, I64, VP, VP, c_i64p, c_i32p, VP, I64, VP]),
    "vb_ingest_max_tiles": (I64, [I64, I64, I64]),
    "vb_ingest_windows": (I64, [I64, I64, I64]),
    "vb_ingest_build": (C.c_int, [VP, VP, VP, VP, VP, VP, VP, I64, C.c_int, VP, VP, I64, I64, I64, I64,
                                  VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, c_i64p, VP, VP,
                                  C.POINTER(VbArrival), I64, C.c_int, VP, I64, VP]),
    "vb_count_components": (C.c_int, [C.POINTER(VbGraph), VP, VP, c_i64p, VP]),
    "vb_offset_copy_i32": (C.c_int, [VP, VP, I64, I32, VP]),
    "vb_add_inplace_f64": (C.c_int, [VP, VP, I64, VP]),
    "vb_gather_stride": (C.c_int, []),
    "vb_pad_blocks": (C.c_int, [VP, VP, I64, VP]),
    "vb_pass_time": (C.c_int, [C.POINTER(VbGraph), C.c_int, VP, VP, VP, VP]),
    "vb_pass_cam": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP]),
    "vb_primal_update": (C.c_int, [VP, VP, VP, VP, I64, VP]),
    "vb_dual_update": (C.c_int, [VP, VP, VP, VP, I64, VP]),
    "vb_gauge_project": (C.c_int, [VP, VP, I64, VP]),
    "vb_so3sync_workspace_bytes": (I64, [I64, I64]),
    "vb_so3sync_run": (C.c_int, [C.POINTER(VbGraph), C.POINTER(VbSo3Options), VP, VP, VP, I64,
                                 C.POINTER(VbSo3Stats), VP]),
    "vb_trans_rhs": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, I64, C.c_int,
                               VP]),
    "vb_trans_cg_workspace_bytes": (I64, [I64, I64]),
    "vb_sell_workspace_bytes": (I64, [I64, I64, I64]),
    "vb_sell_count": (C.c_int, [C.POINTER(VbGraph), VP, VP, c_i64p, c_i64p, VP, I64, VP]),
    "vb_sell_fill": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP, VP, VP, VP, I64, VP, I64, VP]),
    "vb_trans_cg": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP, VP, F64, I64, C.c_int, VP, VP, c_i32p, VP, I64, VP, VP,
                              C.c_int, VP]),
    "vb_trans_lsqr_workspace_bytes": (I64, [I64, I64, I64]),
    "vb_trans_lsqr": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP, VP, VP, VP, I64, VP, VP, F64, F64, F64, I64,
                                c_i32p, c_i32p, VP, I64, VP, VP, VP]),
    "vb_trans_schur_workspace_bytes": (I64, [I64, I64]),
    "vb_trans_schur_direct": (C.c_int, [C.POINTER(VbGraph), VP, VP, VP, VP, VP, I64, VP]),
    "vb_nccl_available": (C.c_int, []),
    "vb_nccl_unique_id": (C.c_int, [VP, I64]),
    "vb_nccl_init": (C.c_int, [VP, I64, C.c_int, C.c_int, C.POINTER(VP)]),
    "vb_nccl_destroy": (C.c_int, [VP]),
    "vb_nccl_allreduce_fn": (VP, []),
    "vb_nccl_allreduce": (C.c_int, [VP, VP, I64, VP]),
    "vb_peer_create": (C.c_int, [C.c_int, C.c_int, I64, C.POINTER(VP), VP]),
    "vb_peer_connect": (C.c_int, [VP, VP]),
    "vb_peer_destroy": (C.c_int, [VP]),
    "vb_peer_allreduce": (C.c_int, [VP, VP, I64, VP]),
    "vb_peer_allreduce_fn": (VP, []),
    "vb_peer_status": (C.c_int, [VP, VP]),
    "vb_peer_stamps": (C.c_int, [VP, VP, VP]),
}


def build(verbose: bool = False) -> str:
    """Compile the extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["bash", BUILD_SCRIPT]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libvican_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    build_flatten()
    return SO_PATH


def flatten_so_path() -> str:
    import sysconfig
    return os.path.join(_HERE, "_vb_flatten" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_flatten() -> str:
    """Compile the host-side dictionary flatten (CPython extension, csrc/flatten.c) in-tree with gcc."""
    import sysconfig
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-Wall", "-I", sysconfig.get_paths()["include"], FLATTEN_SRC,
           "-o", flatten_so_path()]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building _vb_flatten failed:\n" + out.stdout + out.stderr)
    return flatten_so_path()


def load_library():
    """dlopen the extension and bind every symbol (no CUDA call is made)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "vican_b200: CUDA extension %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % SO_PATH)
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def lib():
    """The loaded extension, after checking that a CUDA device is present."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("vican_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load_library()


class VbError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = load_library().vb_status_string(code).decode()
        super().__init__("%s failed with status %d: %s" % (where, code, msg))


def check(code: int, where: str, allow=()):
    if code != 0 and code not in allow:
        raise VbError(code, where)
    return code
