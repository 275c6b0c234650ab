"""Batched fp64 geometry on the device (vican/geometry.py as kernels): SE(3) compose / invert,
SO(3) projection and the SVD factors of the primal / dual updates.  Inputs may be numpy arrays
or torch tensors; outputs are CUDA tensors."""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import check
from .solver import F64, _dev, _ptr, _stream

__all__ = ["se3_compose_batch", "se3_invert_batch", "polar_so3_batch", "svd3_factors_batch", "optimize_gauge_batch",
           "distance_so3_batch", "se3_left_compose_batch"]


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def se3_invert_batch(R, t, round_f32: bool = False):
    """SE3.inv (geometry.py:235-243) for n poses: (R^T, -R^T t); ``round_f32`` mirrors the
    reference's float32 store of the result."""
    lib = _cabi.lib()
    dev = _device()
    R = _dev(R, F64, dev).reshape(-1, 9)
    t = _dev(t, F64, dev).reshape(-1, 3)
    n = R.shape[0]
    Ri, ti = torch.empty_like(R), torch.empty_like(t)
    check(lib.vb_se3_invert_batch(_ptr(R), _ptr(t), _ptr(Ri), _ptr(ti), n, int(round_f32), _stream()),
          "vb_se3_invert_batch")
    return Ri.view(n, 3, 3), ti


def se3_compose_batch(Ra, ta, Rb, tb, round_f32: bool = False):
    """SE3.__matmul__ (geometry.py:260-261) for n pairs: (Ra Rb, Ra tb + ta)."""
    lib = _cabi.lib()
    dev = _device()
    Ra, Rb = _dev(Ra, F64, dev).reshape(-1, 9), _dev(Rb, F64, dev).reshape(-1, 9)
    ta, tb = _dev(ta, F64, dev).reshape(-1, 3), _dev(tb, F64, dev).reshape(-1, 3)
    n = Ra.shape[0]
    Ro, to = torch.empty_like(Ra), torch.empty_like(ta)
    check(lib.vb_se3_compose_batch(_ptr(Ra), _ptr(ta), _ptr(Rb), _ptr(tb), _ptr(Ro), _ptr(to), n, int(round_f32),
                                   _stream()), "vb_se3_compose_batch")
    return Ro.view(n, 3, 3), to


def polar_so3_batch(M):
    """project_SO3 (geometry.py:175-191) for n blocks."""
    lib = _cabi.lib()
    M = _dev(M, F64, _device()).reshape(-1, 9)
    R = torch.empty_like(M)
    check(lib.vb_polar_so3_batch(_ptr(M), _ptr(R), M.shape[0], _stream()), "vb_polar_so3_batch")
    return R.view(-1, 3, 3)


def svd3_factors_batch(M):
    """(rot, (M M^T)^{1/2}, (M M^T)^{-1/2}) per block (bipgo.py:306-312, :323-329)."""
    lib = _cabi.lib()
    M = _dev(M, F64, _device()).reshape(-1, 9)
    a, b, c = torch.empty_like(M), torch.empty_like(M), torch.empty_like(M)
    check(lib.vb_svd3_factors_batch(_ptr(M), _ptr(a), _ptr(b), _ptr(c), M.shape[0], _stream()),
          "vb_svd3_factors_batch")
    return a.view(-1, 3, 3), b.view(-1, 3, 3), c.view(-1, 3, 3)


def optimize_gauge_batch(Ra, ta, Rb, tb):
    """optimize_gauge_SE3 (geometry.py:294-324) on stacked poses: the transform G minimising
    ``a_i - b_i @ G`` over i.  ``ta = tb = None``: optimize_gauge_SO3 (geometry.py:264-291).
    Returns (G_R [3,3], G_t [3] or None) as CUDA tensors."""
    lib = _cabi.lib()
    dev = _device()
    Ra, Rb = _dev(Ra, F64, dev).reshape(-1, 9), _dev(Rb, F64, dev).reshape(-1, 9)
    n = Ra.shape[0]
    if Rb.shape[0] != n:
        raise AssertionError("len(poses_a) == len(poses_b)")          # geometry.py:313
    with_t = ta is not None
    if with_t:
        ta, tb = _dev(ta, F64, dev).reshape(-1, 3), _dev(tb, F64, dev).reshape(-1, 3)
    wsb = int(lib.vb_gauge_workspace_bytes(n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    gR = torch.empty(9, dtype=F64, device=dev)
    gt = torch.empty(3, dtype=F64, device=dev) if with_t else None
    check(lib.vb_optimize_gauge(_ptr(Ra), _ptr(ta) if with_t else None, _ptr(Rb), _ptr(tb) if with_t else None, n,
                                _ptr(gR), _ptr(gt), _ptr(ws), wsb, _stream()), "vb_optimize_gauge")
    return gR.view(3, 3), gt


def distance_so3_batch(R1, R2=None):
    """distance_SO3 (geometry.py:154-172) for n pairs, in degrees; ``R2 = None``: angle(R1)
    (geometry.py:131-151)."""
    lib = _cabi.lib()
    dev = _device()
    R1 = _dev(R1, F64, dev).reshape(-1, 9)
    if R2 is not None:
        R2 = _dev(R2, F64, dev).reshape(-1, 9)
        assert R2.shape == R1.shape
    out = torch.empty(R1.shape[0], dtype=F64, device=dev)
    check(lib.vb_distance_so3_batch(_ptr(R1), _ptr(R2), _ptr(out), R1.shape[0], _stream()), "vb_distance_so3_batch")
    return out


def se3_left_compose_batch(Rg, tg, R, t, round_f32: bool = False):
    """(Rg, tg) @ (R_i, t_i) for one left transform and n poses (main.ipynb cell 9:
    ``G.inv() @ pose_est[c]``)."""
    lib = _cabi.lib()
    dev = _device()
    Rg, tg = _dev(Rg, F64, dev).reshape(9), _dev(tg, F64, dev).reshape(3)
    R, t = _dev(R, F64, dev).reshape(-1, 9), _dev(t, F64, dev).reshape(-1, 3)
    n = R.shape[0]
    Ro, to = torch.empty_like(R), torch.empty_like(t)
    check(lib.vb_se3_left_compose_batch(_ptr(Rg), _ptr(tg), _ptr(R), _ptr(t), _ptr(Ro), _ptr(to), n, int(round_f32),
                                        _stream()), "vb_se3_left_compose_batch")
    return Ro.view(n, 3, 3), to
