"""cfg4-shaped synthetic camera networks generated directly on the device (SURVEY.md 8d:
10 k cameras, 1 M single-marker time nodes, 50 cameras per node -> 50 M edges; the dict API is
impossible at this size).  Generation is blocked by time-node range with one RNG stream per
block, so any rank can generate exactly its shard of the SAME global graph."""
from __future__ import annotations

import dataclasses

import torch

F64 = torch.float64


def _rand_rot(gen, n, device):
    q = torch.randn((n, 4), generator=gen, device=device, dtype=F64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(n, 3, 3)


def _so3_exp(xi):
    th = xi.norm(dim=1).clamp_min(1e-12)
    a = (torch.sin(th) / th)[:, None, None]
    b = ((1 - torch.cos(th)) / (th * th))[:, None, None]
    K = torch.zeros((xi.shape[0], 3, 3), dtype=F64, device=xi.device)
    K[:, 0, 1], K[:, 0, 2] = -xi[:, 2], xi[:, 1]
    K[:, 1, 0], K[:, 1, 2] = xi[:, 2], -xi[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -xi[:, 1], xi[:, 0]
    return torch.eye(3, dtype=F64, device=xi.device)[None] + a * K + b * (K @ K)


@dataclasses.dataclass
class DeviceDetections:
    """Raw detections of one shard, device resident (time indices are LOCAL to the shard)."""
    n_c: int
    n_t: int                 # local time nodes
    t_lo: int                # global index of local node 0
    cam: torch.Tensor        # int32 [E]
    time: torch.Tensor       # int32 [E] local
    marker: torch.Tensor     # int32 [E] (all zero: single-marker nodes, identity constraint)
    R: torch.Tensor          # [E, 9]
    t: torch.Tensor          # [E, 3]
    k_r: torch.Tensor
    k_t: torch.Tensor
    gt_cam_R: torch.Tensor   # [n_c, 3, 3] camera -> world
    gt_time_R: torch.Tensor  # [n_t, 3, 3] local
    gt_cam_t: torch.Tensor
    gt_time_t: torch.Tensor

    @property
    def n_edges(self):
        return int(self.cam.shape[0])


def make_scaled_network(seed: int, n_c: int, n_t: int, cams_per_t: int, t_lo: int = 0, t_hi: int = None,
                        block: int = 125_000, sigma_R: float = 0.02, sigma_t: float = 0.01,
                        device="cuda") -> DeviceDetections:
    """Time nodes [t_lo, t_hi) of the global graph (seed, n_c, n_t, cams_per_t).  Node t is seen
    by ``cams_per_t`` distinct cameras: one uniformly drawn from each of ``cams_per_t`` equal
    camera strata (distinct by construction, so E = E_raw exactly)."""
    t_hi = n_t if t_hi is None else t_hi
    dev = torch.device(device)
    assert n_c % cams_per_t == 0, "n_c must be a multiple of cams_per_t"
    stratum = n_c // cams_per_t
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    cam_R = _rand_rot(gen, n_c, dev)
    cam_t = torch.randn((n_c, 3), generator=gen, device=dev, dtype=F64) * 5.0
    cams, times, Rs, ts, krs, kts, tR, tt = [], [], [], [], [], [], [], []
    b0 = t_lo // block
    b1 = (t_hi + block - 1) // block
    for b in range(b0, b1):
        lo, hi = b * block, min((b + 1) * block, n_t)
        nb = hi - lo
        gen.manual_seed(seed * 1_000_003 + 17 * b + 1)
        Rt = _rand_rot(gen, nb, dev)
        tt_b = torch.randn((nb, 3), generator=gen, device=dev, dtype=F64) * 3.0
        c = torch.randint(0, stratum, (nb, cams_per_t), generator=gen, device=dev) + \
            torch.arange(cams_per_t, device=dev)[None, :] * stratum
        xi = torch.randn((nb * cams_per_t, 3), generator=gen, device=dev, dtype=F64) * sigma_R
        tn = torch.randn((nb * cams_per_t, 3), generator=gen, device=dev, dtype=F64) * sigma_t
        w = torch.rand((nb * cams_per_t,), generator=gen, device=dev, dtype=F64) + 0.5
        # keep only the part of the block that belongs to [t_lo, t_hi)
        s0, s1 = max(t_lo, lo) - lo, min(t_hi, hi) - lo
        sel = slice(s0 * cams_per_t, s1 * cams_per_t)
        cflat = c.reshape(-1)
        tloc = torch.arange(nb, device=dev).repeat_interleave(cams_per_t)
        RcT = cam_R[cflat[sel]].transpose(1, 2)
        Rdet = RcT @ Rt[tloc[sel]] @ _so3_exp(xi[sel])
        tdet = torch.einsum("eij,ej->ei", RcT, tt_b[tloc[sel]] - cam_t[cflat[sel]]) + tn[sel]
        cams.append(cflat[sel].to(torch.int32))
        times.append((tloc[sel] + (lo - t_lo)).to(torch.int32))
        Rs.append(Rdet.reshape(-1, 9)); ts.append(tdet)
        krs.append(w[sel]); kts.append(2.0 * w[sel])
        tR.append(Rt[s0:s1]); tt.append(tt_b[s0:s1])
        del Rt, c, xi, tn, w, RcT, Rdet, tdet, tloc, cflat
    cat = torch.cat
    E_cam = cat(cams)
    return DeviceDetections(n_c, t_hi - t_lo, t_lo, E_cam, cat(times), torch.zeros_like(E_cam), cat(Rs), cat(ts),
                            cat(krs), cat(kts), cam_R, cat(tR), cam_t, cat(tt))
