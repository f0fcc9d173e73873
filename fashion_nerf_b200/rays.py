"""Synthetic pinhole camera of SURVEY.md 8(d): the caller side of render_rays used by the bench and
the full-frame driver (origin (0,0,4), looking down -z, horizontal FOV 0.6911 rad, un-normalised
directions so the |d| factor of A.5 is exercised; view v of V orbits about +y)."""
from __future__ import annotations

import math

import torch


def pinhole_rays(H: int, W: int, view: int = 0, n_views: int = 1, device="cpu"):
    """Returns rays_o, rays_d [H*W, 3] fp32 (computed on the CPU so every consumer sees the same bits)."""
    f = 0.5 * W / math.tan(0.34555)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    d = torch.stack([(i - W * 0.5) / f, -(j - H * 0.5) / f, -torch.ones_like(i)], -1).reshape(-1, 3)
    o = torch.tensor([0.0, 0.0, 4.0]).expand_as(d).clone()
    if n_views > 1 and view != 0:
        a = 2.0 * math.pi * view / n_views
        c, s = math.cos(a), math.sin(a)
        rot = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float32)
        d = d @ rot.t()
        o = o @ rot.t()
    return o.contiguous().to(device), d.contiguous().to(device)
