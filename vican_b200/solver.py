"""Array-level device solver: the path BASELINE.json's north_star names, on raw arrays.

``DeviceGraph`` owns the device-resident block-CSR/CSC built by the ingestion kernels;
``solve_rotations`` runs the primal-dual loop (vican/bipgo.py:145-350) and
``solve_translations`` the least-squares stage (bipgo.py:420-487).  Everything numerical
happens in ``libvican_b200.so``; torch only owns memory and streams.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import VbGraph, VbSo3Options, VbSo3Stats, check

F64 = torch.float64
I32 = torch.int32
# Inexact inner solves (profiles/r2_inexact_inner.md): the eigen-solves of the early outer iterations stop at
# TOL_EARLY instead of 1e-13 when at least MIN_MAXITER iterations are requested; the last EARLY_MARGIN iterations
# are always tight and must find the outer iteration at its fixed point, else the run is repeated all-tight.
TOL_EARLY = float(os.environ.get("VICAN_B200_TOL_EARLY", "1e-5"))
EARLY_MARGIN = 4
EARLY_MIN_MAXITER = 8
SCHUR_MAX_CAMERAS = 4096      # dense direct translation solve: n_c^2 doubles (128 MB at the cap)


class ConvergenceError(AssertionError, RuntimeError):
    """CG did not converge (the reference trips ``assert exit_code == 0``, bipgo.py:478)."""


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device, non_blocking=True)


class HostUpload:
    """Host -> device upload of the raw detection arrays on a side stream, in the order the
    pipeline consumes them (indices first: the key sort starts while the 72-byte rotation blocks
    are still crossing PCIe; translations last: they are only needed after the rotation stage).
    ``get(name)`` makes the CURRENT stream wait for that array's copy and returns the tensor."""

    ORDER = ("cam", "time", "marker", "k_r", "k_t", "R", "t")
    R_CHUNKS = 8       # the rotations (72 of the 92 bytes per detection) cross PCIe in chunks: the fold of chunk k
                       # runs under the copy of chunk k + 1 (vb_arrival)

    def __init__(self, arrays: dict, dtypes: dict, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self._t, self._ev = {}, {}
        cur = torch.cuda.current_stream(device)
        # destinations are allocated on the COMPUTE stream (no cross-stream traffic in the caching
        # allocator: allocating them on the side stream made it fall back to cudaMalloc / cudaFree)
        dst = {}
        for name in self.ORDER:
            if name in arrays:
                src = arrays[name]
                src = src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src))
                arrays[name] = src.contiguous()
                dst[name] = torch.empty(src.shape, dtype=dtypes[name], device=device)
        self.stream.wait_stream(cur)
        self.r_chunks = None
        with torch.cuda.stream(self.stream):
            for name in self.ORDER:
                if name not in dst:
                    continue
                n = dst[name].shape[0]
                if name == "R" and n >= 64 * self.R_CHUNKS:
                    ends, evs = [], []
                    for k in range(self.R_CHUNKS):
                        lo, hi = n * k // self.R_CHUNKS, n * (k + 1) // self.R_CHUNKS
                        dst[name][lo:hi].copy_(arrays[name][lo:hi], non_blocking=True)
                        e = torch.cuda.Event()
                        e.record(self.stream)
                        ends.append(hi); evs.append(e)
                    self.r_chunks = (ends, evs)
                else:
                    dst[name].copy_(arrays[name], non_blocking=True)
                dst[name].record_stream(self.stream)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self._t[name], self._ev[name] = dst[name], ev

    def get(self, name):
        torch.cuda.current_stream(self.device).wait_event(self._ev[name])
        return self._t[name]

    def peek(self, name):
        """The destination tensor WITHOUT waiting for its copy (the consumer synchronises itself)."""
        return self._t[name]


@dataclasses.dataclass
class Comm:
    """Communicator handles of the extension (edge-sharded multi-GPU runs): ``ctx`` = NCCL
    communicator, ``peer`` = NVLink peer-memory windows (``vb_peer_create``) or None.  With
    ``peer`` every cross-rank sum is the one-shot kernel of csrc/peer.cuh and the camera pass runs
    fused with its sum; without it NCCL all-reduces are issued on the solver's stream."""
    ctx: int
    rank: int
    world: int
    peer: Optional[int] = None
    peer_capacity: int = 0

    def reducer(self, lib, count: int):
        """(vb_allreduce_fn, ctx) able to sum ``count`` doubles."""
        if self.peer is not None and count <= self.peer_capacity:
            return lib.vb_peer_allreduce_fn(), self.peer
        return lib.vb_nccl_allreduce_fn(), self.ctx

    def allreduce(self, lib, t: torch.Tensor):
        n = t.numel()
        if self.peer is not None and n <= self.peer_capacity:
            check(lib.vb_peer_allreduce(self.peer, _ptr(t), n, _stream()), "vb_peer_allreduce")
        else:
            check(lib.vb_nccl_allreduce(self.ctx, _ptr(t), n, _stream()), "vb_nccl_allreduce")


def default_tile_len(n_edges: int, n_sms: int = 148) -> int:
    """Camera-tile length: a multiple of 50 edges (one work item of the edge pass), sized so the
    camera pass has a few thousand warps but at most 400 edges per 9 fp64 atomics."""
    tl = (n_edges // (n_sms * 32)) // 50 * 50
    return int(min(max(tl, 50), 400))


class DeviceGraph:
    """Device-resident aggregated bipartite graph (time-sorted CSR + camera-sorted CSC)."""

    def __init__(self, cam, time, marker, R, k_r, k_t, markerC, n_c: int, n_t: int,
                 round_kr_f32: bool = False, device=None, tile_len: Optional[int] = None,
                 upload: Optional[HostUpload] = None):
        lib = _cabi.lib()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self.n_c, self.n_t = int(n_c), int(n_t)
        if upload is not None:        # arrays are crossing PCIe on a side stream: wait only for the indices now
            self.cam, self.time = upload.get("cam"), upload.get("time")
        else:
            self.cam = _dev(cam, I32, dev)
            self.time = _dev(time, I32, dev)
        markerC = _dev(markerC, F64, dev).reshape(-1, 9)
        n_raw = int(self.cam.shape[0])
        if n_raw == 0:
            raise ValueError("no edges survive edge_filter")
        self.n_raw = n_raw
        with torch.cuda.device(dev):
            wsb = int(lib.vb_ingest_workspace_bytes(n_raw))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            self.raw_perm = torch.empty(n_raw, dtype=I32, device=dev)
            self.raw_pair = torch.empty(n_raw, dtype=I32, device=dev)
            npairs, was_sorted = C.c_int64(0), C.c_int32(0)
            check(lib.vb_ingest_sort(_ptr(self.cam), _ptr(self.time), n_raw, self.n_c, self.n_t, _ptr(self.raw_perm),
                                     _ptr(self.raw_pair), C.byref(npairs), C.byref(was_sorted), _ptr(ws), wsb, _stream()),
                  "vb_ingest_sort")
            E = int(npairs.value)
            self.n_edges = E
            self.raw_sorted = bool(was_sorted.value)      # raw_perm is the identity
            arrival = None
            if upload is not None:
                self.marker, self.k_r, self.k_t = upload.get("marker"), upload.get("k_r"), upload.get("k_t")
                if upload.r_chunks is not None:
                    # the rotations are still in flight: hand the extension their arrival schedule instead of waiting
                    ends, evs = upload.r_chunks
                    self._arr_ends = (C.c_int64 * len(ends))(*ends)
                    self._arr_evs = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
                    arrival = _cabi.VbArrival(len(ends), int(was_sorted.value), self._arr_ends, self._arr_evs)
                    R = upload.peek("R").reshape(-1, 9)
                else:
                    R = upload.get("R").reshape(-1, 9)
            else:
                self.marker = _dev(marker, I32, dev)
                R = _dev(R, F64, dev).reshape(-1, 9)
                self.k_r = _dev(k_r, F64, dev)
                self.k_t = _dev(k_t, F64, dev)
            tl = default_tile_len(E) if tile_len is None else int(tile_len)
            self.tile_len = tl
            max_tiles = int(lib.vb_ingest_max_tiles(E, self.n_c, tl))
            e = lambda n, dt: torch.empty(n, dtype=dt, device=dev)  # noqa: E731
            # the edge passes stream blocks / indices with 16-byte granular bulk copies: 2 blocks / 8 indices
            # of padding behind the arrays keep those reads in bounds.  Only the tail is cleared (the padding is
            # staged but never consumed); zero-filling the whole arrays cost 7.6 GB of writes per ingestion.
            self._t_cam_pad, self._t_B_pad = e(E + 8, I32), e((E + 2, 9), F64)
            self._c_time_pad, self._c_B_pad = e(E + 8, I32), e((E + 2, 9), F64)
            for pad_arr in (self._t_cam_pad, self._c_time_pad):
                pad_arr[E:].zero_()
            for pad_arr in (self._t_B_pad, self._c_B_pad):
                pad_arr[E:].zero_()
            self.t_cam, self.t_B = self._t_cam_pad[:E], self._t_B_pad[:E]
            self.c_time, self.c_B = self._c_time_pad[:E], self._c_B_pad[:E]
            self.t_rowptr, self.t_time = e(self.n_t + 1, I32), e(E, I32)
            self.t_a, self.t_w = e(E, F64), e(E, F64)
            self.pair_start = e(E + 1, I32)
            self.n_windows = int(lib.vb_ingest_windows(E, self.n_c, tl))
            if self.n_windows * self.n_c + 1 > n_raw:     # tiny chunk: fewer detections than (window, camera) runs
                wsb = int(lib.vb_ingest_workspace_bytes(self.n_windows * self.n_c + 1))
                ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            self.c_segptr = e(self.n_windows * self.n_c + 1, I32)
            self.c_w, self.c_order = e(E, F64), e(E, I32)
            self.tile_cam, self.tile_start = e(max_tiles + 1, I32), e(max_tiles + 1, I32)
            self.tile_off = e(self.n_windows * self.n_c + 1, I32)
            self.deg_t, self.deg_c = e(self.n_t, F64), e(self.n_c, F64)
            ntiles = C.c_int64(0)
            check(lib.vb_ingest_build(
                _ptr(self.cam), _ptr(self.time), _ptr(self.marker), _ptr(R), _ptr(self.k_r), _ptr(self.k_t),
                _ptr(markerC), n_raw, 1 if round_kr_f32 else 0, _ptr(self.raw_perm), _ptr(self.raw_pair), E,
                self.n_c, self.n_t, tl, _ptr(self.t_rowptr), _ptr(self.t_cam), _ptr(self.t_time), _ptr(self.t_B),
                _ptr(self.t_a), _ptr(self.t_w), _ptr(self.pair_start), _ptr(self.c_segptr), _ptr(self.c_time),
                _ptr(self.c_B), _ptr(self.c_w), _ptr(self.c_order), _ptr(self.tile_cam),
                _ptr(self.tile_start),
                _ptr(self.tile_off), C.byref(ntiles), _ptr(self.deg_t), _ptr(self.deg_c),
                C.byref(arrival) if arrival is not None else None, int(markerC.shape[0]),
                1 if self.raw_sorted else 0, _ptr(ws), wsb, _stream()),
                "vb_ingest_build")
            self.n_tiles = int(ntiles.value)
            self.tile_part = e((max(self.n_tiles, 1), 9), F64)      # camera-pass scratch (per-tile sums)
        del ws, R
        self.cgraph = VbGraph(
            self.n_c, self.n_t, E, self.n_tiles, self.n_windows,
            self.t_rowptr.data_ptr(), self.t_cam.data_ptr(), self.t_B.data_ptr(), self.t_w.data_ptr(),
            self.c_segptr.data_ptr(), self.c_order.data_ptr(), self.c_time.data_ptr(), self.c_B.data_ptr(),
            self.c_w.data_ptr(),
            self.tile_cam.data_ptr(), self.tile_start.data_ptr(), self.tile_off.data_ptr(), self.tile_part.data_ptr(),
            self.deg_t.data_ptr(), self.deg_c.data_ptr(), 0, 0, 0, 0, 0, 0)
        self._sell = None

    def ensure_sell(self):
        """Sliced-ELL copy of the translation Laplacian (both sides), built on first use by the
        conjugate-gradient solver (csrc/cg.cuh): 12 bytes per stored slot and side."""
        if self._sell is not None:
            return
        lib = _cabi.lib()
        dev = self.device
        with torch.cuda.device(dev):
            ns_t, ns_c = (self.n_t + 7) // 8, (self.n_c + 7) // 8
            st_ptr = torch.empty(ns_t + 2, dtype=I32, device=dev)
            sc_ptr = torch.empty(ns_c + 2, dtype=I32, device=dev)
            wsb = int(lib.vb_sell_workspace_bytes(self.n_c, self.n_t, self.n_windows))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            ct, cc = C.c_int64(0), C.c_int64(0)
            check(lib.vb_sell_count(C.byref(self.cgraph), _ptr(st_ptr), _ptr(sc_ptr), C.byref(ct), C.byref(cc),
                                    _ptr(ws), wsb, _stream()), "vb_sell_count")
            st_idx = torch.empty(32 * max(int(ct.value), 1), dtype=I32, device=dev)
            st_w = torch.empty(32 * max(int(ct.value), 1), dtype=F64, device=dev)
            sc_idx = torch.empty(32 * max(int(cc.value), 1), dtype=I32, device=dev)
            sc_w = torch.empty(32 * max(int(cc.value), 1), dtype=F64, device=dev)
            check(lib.vb_sell_fill(C.byref(self.cgraph), _ptr(st_ptr), _ptr(st_idx), _ptr(st_w), _ptr(sc_ptr),
                                   _ptr(sc_idx), _ptr(sc_w), int(cc.value), _ptr(ws), wsb, _stream()), "vb_sell_fill")
        self._sell = (st_ptr, st_idx, st_w, sc_ptr, sc_idx, sc_w)
        self.sell_chunks = (int(ct.value), int(cc.value))
        g = self.cgraph
        g.st_ptr, g.st_idx, g.st_w = st_ptr.data_ptr(), st_idx.data_ptr(), st_w.data_ptr()
        g.sc_ptr, g.sc_idx, g.sc_w = sc_ptr.data_ptr(), sc_idx.data_ptr(), sc_w.data_ptr()

    def n_components(self) -> int:
        """Connected components of the aggregated bipartite graph (device: hooking + pointer jumping)."""
        lib = _cabi.lib()
        with torch.cuda.device(self.device):
            labels = torch.empty(self.n_c + self.n_t + 2, dtype=I32, device=self.device)
            cnt = C.c_int64(0)
            check(lib.vb_count_components(C.byref(self.cgraph), _ptr(self.t_time), _ptr(labels), C.byref(cnt), _stream()),
                  "vb_count_components")
        return int(cnt.value)

    # ---- algorithmic bytes of one edge pass (76 B / edge + node traffic) ----
    def pass_bytes(self, kind: str) -> int:
        E, n_c, n_t = self.n_edges, self.n_c, self.n_t
        if kind == "time":     # blocks + cam index, row pointers, Lambda_T read, W write (padded), X gather source
            return 76 * E + 4 * (n_t + 1) + 72 * n_t + 96 * n_t + 96 * n_c
        if kind == "cam":      # blocks + time index, tile table, W gather source (padded), Y accumulate
            return 76 * E + (8 + 2 * 72) * self.n_tiles + 96 * n_t + 72 * n_c
        raise ValueError(kind)


class StreamingGraph:
    """Device graph that grows by ranges of NEW time nodes (the caller side of the path,
    ``estimate_pose_mp`` output arriving image by image, cam.py:176-185, :243-263).

    ``append`` ingests one chunk on its own (same kernels as ``DeviceGraph``, chunk-local time indices)
    and places its arrays behind the resident ones: the time-sorted block-CSR is append-only in time, and
    the chunk's time windows become new windows of the camera-pass order, so nothing already on the device
    is re-sorted or rewritten -- only index offsets are added (``vb_offset_copy_i32``) and the camera
    degrees accumulated (``vb_add_inplace_f64``).  Cost per append is proportional to the chunk.
    ``graph()`` returns a ``DeviceGraph`` view of the current state for ``solve_rotations`` /
    ``solve_translations``; ``t`` are the detections' translations in raw order.  Storage grows by
    doubling.  A chunk should hold a few thousand detections (one chunk = at least one camera-pass
    window: ``io.EdgeAccumulator`` does the host-side batching)."""

    _I32 = ("t_cam", "t_time", "t_rowptr", "pair_start", "c_time", "c_order", "c_segptr", "tile_cam", "tile_start",
            "tile_off", "cam", "time", "marker", "raw_perm", "raw_pair")

    def __init__(self, n_c: int, markerC, device=None, round_kr_f32: bool = False, capacity_edges: int = 1 << 16):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_c = int(n_c)
        self.markerC = _dev(markerC, F64, self.device).reshape(-1, 9)
        self.round_kr_f32 = bool(round_kr_f32)
        self.n_t = self.n_edges = self.n_raw = self.n_tiles = self.n_windows = 0
        self.raw_sorted = True
        self._cap0 = int(capacity_edges)
        self._a = {}
        z = lambda n, dt, w=None: torch.zeros((n,) if w is None else (n, w), dtype=dt, device=self.device)  # noqa: E731
        self._a["deg_c"] = z(self.n_c, F64)
        for name in self._I32:
            self._a[name] = z(1, I32)
        for name, w in (("t_B", 9), ("c_B", 9), ("t_a", None), ("t_w", None), ("c_w", None), ("deg_t", None),
                        ("k_r", None), ("k_t", None), ("t", 3)):
            self._a[name] = z(1, F64, w)

    def _need(self, name, n):
        """Make array ``name`` hold at least ``n`` leading entries (amortised doubling, old content kept)."""
        a = self._a[name]
        if a.shape[0] >= n:
            return a
        cap = max(n, 2 * a.shape[0], self._cap0)
        b = torch.zeros((cap,) + tuple(a.shape[1:]), dtype=a.dtype, device=self.device)
        b[:a.shape[0]] = a
        self._a[name] = b
        return b

    def append(self, cam, time_local, marker, R, t, k_r, k_t, n_t_chunk: int) -> None:
        """Detections of ``n_t_chunk`` new time nodes; ``time_local`` counts from 0 inside the chunk (global
        index = current ``n_t`` + local)."""
        lib = _cabi.lib()
        c = DeviceGraph(cam, time_local, marker, R, k_r, k_t, self.markerC, self.n_c, int(n_t_chunk),
                        round_kr_f32=self.round_kr_f32, device=self.device)
        E0, N0, R0, T0, W0 = self.n_edges, self.n_t, self.n_raw, self.n_tiles, self.n_windows
        Ec, Nc, Rc, Tc, Wc, n_c = c.n_edges, c.n_t, c.n_raw, c.n_tiles, c.n_windows, self.n_c
        self.raw_sorted = self.raw_sorted and c.raw_sorted
        with torch.cuda.device(self.device):
            def put(name, src, lo):                       # plain device-to-device append
                self._need(name, lo + src.shape[0] + (8 if name in ("t_cam", "c_time") else 2 if name in ("t_B", "c_B") else 0))
                self._a[name][lo:lo + src.shape[0]] = src

            def put_off(name, src, lo, add):              # append with an index offset (one small kernel)
                dst = self._need(name, lo + src.shape[0] + (8 if name in ("t_cam", "c_time") else 0))
                check(lib.vb_offset_copy_i32(C.c_void_p(dst.data_ptr() + 4 * lo), _ptr(src), src.shape[0], int(add), _stream()),
                      "vb_offset_copy_i32")
            put("t_B", c.t_B, E0); put("c_B", c.c_B, E0)
            put("t_a", c.t_a, E0); put("t_w", c.t_w, E0); put("c_w", c.c_w, E0)
            put("deg_t", c.deg_t, N0)
            put("k_r", c.k_r, R0); put("k_t", c.k_t, R0); put("t", _dev(t, F64, self.device).reshape(-1, 3), R0)
            put("cam", c.cam, R0); put("marker", c.marker, R0); put("t_cam", c.t_cam, E0)
            put("tile_cam", c.tile_cam[:Tc], T0)
            put_off("t_time", c.t_time, E0, N0)
            put_off("t_rowptr", c.t_rowptr, N0, E0)
            put_off("pair_start", c.pair_start, E0, R0)
            put_off("c_time", c.c_time, E0, N0)
            put_off("c_order", c.c_order, E0, E0)
            put_off("c_segptr", c.c_segptr, W0 * n_c, E0)
            put_off("tile_start", c.tile_start[:Tc + 1], T0, E0)
            put_off("tile_off", c.tile_off, W0 * n_c, T0)
            put_off("time", c.time, R0, N0)
            put_off("raw_perm", c.raw_perm, R0, R0)
            put_off("raw_pair", c.raw_pair, R0, E0)
            check(lib.vb_add_inplace_f64(_ptr(self._a["deg_c"]), _ptr(c.deg_c), n_c, _stream()), "vb_add_inplace_f64")
        self.n_edges, self.n_t, self.n_raw, self.n_tiles, self.n_windows = E0 + Ec, N0 + Nc, R0 + Rc, T0 + Tc, W0 + Wc

    @property
    def t(self) -> torch.Tensor:
        return self._a["t"][:self.n_raw]

    def graph(self) -> "DeviceGraph":
        """``DeviceGraph`` view of the current state (shares storage with this object)."""
        g = DeviceGraph.__new__(DeviceGraph)
        a, E, N, Rn, T, W = self._a, self.n_edges, self.n_t, self.n_raw, self.n_tiles, self.n_windows
        if E == 0:
            raise ValueError("no detections appended yet")
        g.device, g.n_c, g.n_t, g.n_edges, g.n_raw, g.n_tiles, g.n_windows = self.device, self.n_c, N, E, Rn, T, W
        g.tile_len = None
        g.raw_sorted = self.raw_sorted
        self._need("t_cam", E + 8); self._need("c_time", E + 8); self._need("t_B", E + 2); self._need("c_B", E + 2)
        a = self._a
        for name, n in (("t_cam", E), ("t_time", E), ("t_B", E), ("t_a", E), ("t_w", E), ("c_time", E), ("c_B", E), ("c_w", E),
                        ("c_order", E), ("t_rowptr", N + 1), ("pair_start", E + 1), ("c_segptr", W * self.n_c + 1),
                        ("tile_cam", T), ("tile_start", T + 1), ("tile_off", W * self.n_c + 1), ("deg_t", N),
                        ("cam", Rn), ("time", Rn), ("marker", Rn), ("k_r", Rn), ("k_t", Rn), ("raw_perm", Rn), ("raw_pair", Rn)):
            setattr(g, name, a[name][:n])
        g.deg_c = a["deg_c"]
        g.tile_part = torch.empty((max(T, 1), 9), dtype=F64, device=self.device)
        g.cgraph = VbGraph(
            g.n_c, N, E, T, W, g.t_rowptr.data_ptr(), g.t_cam.data_ptr(), g.t_B.data_ptr(), g.t_w.data_ptr(),
            g.c_segptr.data_ptr(), g.c_order.data_ptr(), g.c_time.data_ptr(), g.c_B.data_ptr(), g.c_w.data_ptr(),
            g.tile_cam.data_ptr(), g.tile_start.data_ptr(), g.tile_off.data_ptr(), g.tile_part.data_ptr(),
            g.deg_t.data_ptr(), g.deg_c.data_ptr(), 0, 0, 0, 0, 0, 0)
        g._sell = None
        return g


@dataclasses.dataclass
class RotationResult:
    r_c: torch.Tensor           # [n_c, 9] as stored by the reference before its final transpose
    r_t: torch.Tensor           # [n_t, 9]
    stats: VbSo3Stats
    status: int

    def world_rotations(self):
        """bipgo.py:344-348: out = r^T."""
        return (self.r_c.view(-1, 3, 3).transpose(1, 2).contiguous(),
                self.r_t.view(-1, 3, 3).transpose(1, 2).contiguous())


def solve_rotations(g: DeviceGraph, maxiter: int, tol: float = 1e-13, max_inner: int = 200,
                    comm: Optional[Comm] = None, profile_events: bool = False, shortcut: bool = True,
                    spanning_start: bool = True, eval_gap: bool = False, tol_early: Optional[float] = None,
                    early_margin: int = EARLY_MARGIN) -> RotationResult:
    """``shortcut=False`` forces the primal multiply through its two edge passes in every outer
    iteration (see ``vb_so3_stats.shortcut_outer``); ``spanning_start=False`` starts the first eigen-solve
    from identity blocks instead of the one-hop estimate around the gauge camera.  Both only change the
    work done, not the result (agreement to rounding / to the eigen-solver's tolerance).
    ``eval_gap=True`` also computes the reference's diagnostics in every outer iteration (the five
    eigenvalues nearest zero, ``stats.evals_hist``) and applies its early exit ``max |lambda_1..5| <=
    1e-6`` (bipgo.py:283-292), at the price of a second eigen-solve per iteration.
    ``tol_early``: eigen tolerance of the outer iterations followed by at least ``early_margin`` more (default:
    ``TOL_EARLY`` when ``maxiter >= EARLY_MIN_MAXITER``, else off).  The primal-dual iteration contracts so
    strongly that the early eigen-solves need not be exact: the tight end-game reaches the same fixed point (measured
    deviation from the all-tight run: 3e-16 rad).  If the last two tight iterations do not accept their start block
    at the first step, the history still matters and the run is repeated with ``tol`` everywhere."""
    lib = _cabi.lib()
    dev = g.device
    if g.n_c < 3:
        raise ValueError("the rotation stage needs at least 3 camera nodes (got %d)" % g.n_c)
    with torch.cuda.device(dev):
        wsb = int(lib.vb_so3sync_workspace_bytes(g.n_c, g.n_t))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        r_c = torch.empty((g.n_c, 9), dtype=F64, device=dev)
        r_t = torch.empty((g.n_t, 9), dtype=F64, device=dev)
        fn, fctx = comm.reducer(lib, 9 * g.n_c) if comm is not None else (None, None)
        fused = comm.peer if (comm is not None and comm.peer is not None and 9 * g.n_c <= comm.peer_capacity) else None
        if tol_early is None:
            tol_early = TOL_EARLY if (maxiter >= EARLY_MIN_MAXITER and not eval_gap) else 0.0
        for attempt in range(2):
            opt = VbSo3Options(int(maxiter), int(max_inner), float(tol), fn, fctx, 1 if profile_events else 0,
                               0 if shortcut else 1, 0 if spanning_start else 1, 1 if eval_gap else 0,
                               float(tol_early), int(early_margin), 0, fused)
            stats = VbSo3Stats()
            rc = lib.vb_so3sync_run(C.byref(g.cgraph), C.byref(opt), _ptr(r_c), _ptr(r_t), _ptr(ws), wsb,
                                    C.byref(stats), _stream())
            check(rc, "vb_so3sync_run", allow=(2,))
            if not stats.inexact_unverified:
                break
            tol_early = 0.0        # the tight end-game did not find a fixed point: repeat with exact inner solves
    res = RotationResult(r_c, r_t, stats, rc)
    res.repeated_tight = bool(attempt == 1)
    return res


@dataclasses.dataclass
class TranslationResult:
    x_c: torch.Tensor           # [n_c, 3]
    x_t: torch.Tensor           # [n_t, 3]
    iters: int
    istop: int


def solve_translations(g: DeviceGraph, rot: RotationResult, t_cm, marker_q, lsqr_solver: str,
                       mode: str = "parity", comm: Optional[Comm] = None, unknown_index=None) -> TranslationResult:
    """``lsqr_solver``: "conjugate_gradient" | "direct" (reference names, bipgo.py:476-480).
    ``mode``: "parity" replays scipy's truncated iterations (what the reference returns);
    "accurate" returns the minimum-norm minimiser itself (up to ~4e-3 away from the reference's
    truncated CG answer -- SURVEY.md 7.3-1): Jacobi-preconditioned CG to 1e-12 for
    "conjugate_gradient", a dense Cholesky of the camera Schur complement for "direct"
    (single GPU, n_c <= SCHUR_MAX_CAMERAS; larger or sharded graphs use the PCG).
    ``unknown_index`` = (unk_c [n_c], unk_t [n_t]): position of every camera / time node in the
    reference's unknown vector (bipgo.py:420-430); it fixes where the diagonal entry sits in the
    row sums of the replayed CSR product.  None: cameras first, then time nodes."""
    if lsqr_solver not in ("conjugate_gradient", "direct"):
        raise ValueError("lsqr_solver must be 'conjugate_gradient' or 'direct', got %r" % (lsqr_solver,))
    lib = _cabi.lib()
    dev = g.device
    with torch.cuda.device(dev):
        t_cm = _dev(t_cm, F64, dev).reshape(-1, 3)
        marker_q = _dev(marker_q, F64, dev).reshape(-1, 3)
        E = g.n_edges
        pair_g = torch.empty((E, 3), dtype=F64, device=dev)
        need_rows = (lsqr_solver == "direct" and mode == "parity")
        d_sorted = torch.empty((g.n_raw, 3), dtype=F64, device=dev) if need_rows else None
        rhs_c = torch.empty((g.n_c, 3), dtype=F64, device=dev)
        rhs_t = torch.empty((g.n_t, 3), dtype=F64, device=dev)
        r_c_pad = torch.empty((g.n_c, int(lib.vb_gather_stride())), dtype=F64, device=dev)
        check(lib.vb_trans_rhs(C.byref(g.cgraph), _ptr(g.raw_perm), _ptr(g.pair_start), _ptr(g.marker), _ptr(t_cm),
                               _ptr(g.k_t), _ptr(marker_q), _ptr(rot.r_c), _ptr(rot.r_t), _ptr(g.t_time),
                               _ptr(pair_g), _ptr(d_sorted), _ptr(rhs_c), _ptr(rhs_t), _ptr(r_c_pad),
                               int(marker_q.shape[0]), 1 if getattr(g, "raw_sorted", False) else 0, _stream()),
              "vb_trans_rhs")
        x_c = torch.empty((g.n_c, 3), dtype=F64, device=dev)
        x_t = torch.empty((g.n_t, 3), dtype=F64, device=dev)
        iters = C.c_int32(0)
        istop = C.c_int32(0)
        if lsqr_solver == "direct" and mode == "accurate" and comm is None and g.n_c <= SCHUR_MAX_CAMERAS:
            # dense block path (north_star (3)): closed-form elimination of the time nodes + Cholesky of the
            # n_c x n_c camera Schur complement, written in the extension (no cuSOLVER)
            wsb = int(lib.vb_trans_schur_workspace_bytes(g.n_c, g.n_t))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            check(lib.vb_trans_schur_direct(C.byref(g.cgraph), _ptr(rhs_c), _ptr(rhs_t), _ptr(x_c), _ptr(x_t), _ptr(ws),
                                            wsb, _stream()), "vb_trans_schur_direct")
            return TranslationResult(x_c, x_t, 0, 0)
        if need_rows:
            wsb = int(lib.vb_trans_lsqr_workspace_bytes(g.n_c, g.n_t, g.n_raw))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            n_unknowns = 3 * (g.n_c + g.n_t)
            fn, fctx = comm.reducer(lib, 3 * g.n_c) if comm is not None else (None, None)
            check(lib.vb_trans_lsqr(C.byref(g.cgraph), _ptr(g.raw_perm), _ptr(g.raw_pair), _ptr(g.pair_start),
                                    _ptr(g.t_time), _ptr(g.k_t), _ptr(d_sorted), g.n_raw, _ptr(x_c), _ptr(x_t),
                                    1e-6, 1e-6, 1e8, 2 * n_unknowns, C.byref(istop), C.byref(iters), _ptr(ws), wsb,
                                    fn, fctx, _stream()), "vb_trans_lsqr")
        else:
            if comm is not None and rhs_c is not None:
                # camera rows of J^T t~ are partial sums over the local edge shard
                comm.allreduce(lib, rhs_c)
            wsb = int(lib.vb_trans_cg_workspace_bytes(g.n_c, g.n_t))
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            n_unknowns = 3 * (g.n_c + g.n_t)   # global count differs on shards; only the cap depends on it
            jacobi = 1 if mode == "accurate" else 0
            rtol = 1e-12 if mode == "accurate" else 1e-5
            fn, fctx = comm.reducer(lib, 3 * g.n_c + 8) if comm is not None else (None, None)
            g.ensure_sell()
            unk_c = unk_t = None
            if unknown_index is not None:
                unk_c, unk_t = (_dev(u, I32, dev) for u in unknown_index)
            rc = lib.vb_trans_cg(C.byref(g.cgraph), _ptr(rhs_c), _ptr(rhs_t), _ptr(x_c), _ptr(x_t), rtol,
                                 10 * n_unknowns, jacobi, _ptr(unk_c), _ptr(unk_t), C.byref(iters), _ptr(ws), wsb,
                                 fn, fctx, 1 if (comm is None or comm.rank == 0) else 0, _stream())
            if rc == 1:
                raise ConvergenceError("conjugate gradient did not converge (reference: assert exit_code == 0)")
            check(rc, "vb_trans_cg")
            if jacobi:
                # J^T J is singular (global translation).  Plain CG from x0 = 0 stays orthogonal to the
                # null space (minimum-norm solution, what the reference returns); the preconditioned
                # iteration does not, so remove the mean explicitly (3 numbers; plumbing, not a kernel).
                tot = torch.cat([x_c.sum(0) , x_t.sum(0), torch.tensor([float(g.n_t)], dtype=F64, device=dev)])
                if comm is not None:
                    tot[:3] = 0.0  # camera block is replicated: count it once, below
                    comm.allreduce(lib, tot)
                    tot[:3] = x_c.sum(0)
                mean = (tot[:3] + tot[3:6]) / (g.n_c + tot[6])
                x_c -= mean
                x_t -= mean
    return TranslationResult(x_c, x_t, int(iters.value), int(istop.value))


_PINNED_OUT = {}
_DOWNLOAD_STREAM = {}


def _to_pinned_host(name: str, t: torch.Tensor, reuse: bool) -> torch.Tensor:
    """Device -> host copy into a PINNED buffer (pageable destinations run at a few GB/s).  With
    ``reuse`` the buffer is cached by (name, shape) and OVERWRITTEN by the next call with the same
    shapes (results of an earlier call alias it); otherwise every call gets fresh buffers."""
    if not reuse:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t, non_blocking=True)
        return buf
    key = (name, tuple(t.shape), t.dtype)
    buf = _PINNED_OUT.get(key)
    if buf is None:
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        _PINNED_OUT[key] = buf
    buf.copy_(t, non_blocking=True)
    return buf


@dataclasses.dataclass
class SolveResult:
    Rw_c: torch.Tensor          # [n_c, 3, 3] world rotations of the cameras
    Rw_t: torch.Tensor          # [n_t, 3, 3] world rotations of the (local) time nodes
    x_c: torch.Tensor           # [n_c, 3]
    x_t: torch.Tensor           # [n_t, 3]
    graph: DeviceGraph
    rot: RotationResult
    trans: TranslationResult
    phase_ms: dict


def solve_arrays(cam, time, marker, R, t, k_r, k_t, markerC, marker_q, n_c: int, n_t: int, maxiter: int,
                 lsqr_solver: str = "conjugate_gradient", mode: str = "parity", tol: float = 1e-13,
                 comm: Optional[Comm] = None, round_kr_f32: bool = False, to_host: bool = False,
                 graph: Optional[DeviceGraph] = None, profile_events: bool = False,
                 reuse_host_buffers: bool = False) -> SolveResult:
    """Array fast path of ``bipartite_se3sync`` (no dicts, no Python callables): raw detections
    as arrays (numpy / pinned host tensors / CUDA tensors) with pre-evaluated weights
    ``k_r = noise_model_r(e)``, ``k_t = noise_model_t(e)`` and already filtered by
    ``edge_filter``; node indices dense and in the caller's order (index 0 = gauge camera).
    With ``to_host`` the results are copied back to pinned host memory (``reuse_host_buffers``: into
    buffers cached across calls -- a later call with the same shapes overwrites them)."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    upload = None
    if graph is None and isinstance(R, torch.Tensor) and not R.is_cuda:
        dev = torch.device("cuda", torch.cuda.current_device())
        upload = HostUpload(dict(cam=cam, time=time, marker=marker, k_r=k_r, k_t=k_t, R=R, t=t),
                            dict(cam=I32, time=I32, marker=I32, k_r=F64, k_t=F64, R=F64, t=F64), dev)
    g = graph if graph is not None else DeviceGraph(cam, time, marker, R, k_r, k_t, markerC, n_c, n_t,
                                                     round_kr_f32=round_kr_f32, upload=upload)
    ev[1].record()
    rot = solve_rotations(g, maxiter, tol=tol, comm=comm, profile_events=profile_events)
    ev[2].record()
    Rw_c, Rw_t = rot.world_rotations()
    if to_host:
        # the rotations (72 of the 96 result bytes per node) go home on a side stream under the translation stage
        cur = torch.cuda.current_stream(g.device)
        dl = _DOWNLOAD_STREAM.get(g.device)
        if dl is None:
            dl = _DOWNLOAD_STREAM[g.device] = torch.cuda.Stream(device=g.device)
        dl.wait_stream(cur)
        with torch.cuda.stream(dl):
            host_R = [_to_pinned_host(n, v, reuse_host_buffers) for n, v in (("Rw_c", Rw_c), ("Rw_t", Rw_t))]
        for v in (Rw_c, Rw_t):
            v.record_stream(dl)
    tr = solve_translations(g, rot, upload.get("t") if upload is not None else t, marker_q, lsqr_solver,
                            mode=mode, comm=comm)
    x_c, x_t = tr.x_c, tr.x_t
    if to_host:
        x_c, x_t = (_to_pinned_host(n, v, reuse_host_buffers) for n, v in (("x_c", x_c), ("x_t", x_t)))
        cur.wait_stream(dl)
        Rw_c, Rw_t = host_R
    ev[3].record()
    torch.cuda.synchronize()
    if comm is not None and comm.peer is not None:
        check(_cabi.lib().vb_peer_status(comm.peer, _stream()), "vb_peer_status")
    phase = dict(ingest=ev[0].elapsed_time(ev[1]), rotation=ev[1].elapsed_time(ev[2]),
                 translation=ev[2].elapsed_time(ev[3]))
    return SolveResult(Rw_c, Rw_t, x_c, x_t, g, rot, tr, phase)
