"""Batched fp64 geometry on the device (vican/geometry.py as kernels): SE(3) compose / invert,
SO(3) projection and the SVD factors of the primal / dual updates.  Inputs may be numpy arrays
or torch tensors; outputs are CUDA tensors."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import check
from .solver import F64, _dev, _ptr, _stream

__all__ = ["se3_compose_batch", "se3_invert_batch", "polar_so3_batch", "svd3_factors_batch"]


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
