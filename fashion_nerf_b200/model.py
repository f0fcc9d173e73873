"""Network parameters of the render path: fp32 master copies in the flat nn.Linear layout of
include/fnerf.h plus the packed device blobs the kernels read."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops

WIDTH, DEPTH, SKIP = 256, 8, 4
PE_XYZ, PE_DIR, COND_DIM = 63, 27, 256

LAYER_NAMES = [f"pts_linears.{i}" for i in range(DEPTH)] + [
    "alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"]


def layer_shapes(cond: bool = False) -> Dict[str, tuple]:
    """(out, in) per nn.Linear, in flat order (nerf-pytorch parameter names)."""
    s = {}
    for i in range(DEPTH):
        k = PE_XYZ if i == 0 else (PE_XYZ + (COND_DIM if cond else 0) + WIDTH if i == SKIP + 1 else WIDTH)
        s[f"pts_linears.{i}"] = (WIDTH, k)
    s["alpha_linear"] = (1, WIDTH)
    s["feature_linear"] = (WIDTH, WIDTH)
    s["views_linears.0"] = (WIDTH // 2, WIDTH + PE_DIR)
    s["rgb_linear"] = (3, WIDTH // 2)
    return s


def init_state_dict(seed: int, cond: bool = False) -> Dict[str, torch.Tensor]:
    """nn.Linear default init, U(-1/sqrt(in), 1/sqrt(in)) for weight and bias, from a private CPU
    generator (seed 0 = coarse, 1 = fine)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, (o, i) in layer_shapes(cond).items():
        bound = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    return sd


def flatten_state_dict(sd: Dict[str, torch.Tensor], cond: bool = False) -> torch.Tensor:
    parts = []
    for name, (o, i) in layer_shapes(cond).items():
        w, b = sd[name + ".weight"], sd[name + ".bias"]
        assert tuple(w.shape) == (o, i) and tuple(b.shape) == (o,), (name, w.shape, b.shape)
        parts += [w.reshape(-1).float(), b.reshape(-1).float()]
    return torch.cat(parts)


def unflatten(flat: torch.Tensor, cond: bool = False) -> Dict[str, torch.Tensor]:
    sd, off = {}, 0
    for name, (o, i) in layer_shapes(cond).items():
        sd[name + ".weight"] = flat[off:off + o * i].reshape(o, i)
        off += o * i
        sd[name + ".bias"] = flat[off:off + o]
        off += o
    assert off == flat.numel()
    return sd


class NerfNetwork:
    """One 8x256 network: ``flat`` (fp32 master, requires_grad for training) + ``packed`` blob."""

    def __init__(self, flat: torch.Tensor, cond: bool = False):
        if not flat.is_cuda:
            raise RuntimeError("NerfNetwork lives on a CUDA device (no CPU fallback)")
        self.cond = cond
        self.flat = flat.detach().clone().float().contiguous()
        assert self.flat.numel() == ops.param_count(cond)
        self.packed = ops.pack_weights(self.flat, cond)

    @classmethod
    def from_state_dict(cls, sd, device, cond: bool = False) -> "NerfNetwork":
        return cls(flatten_state_dict(sd, cond).to(device), cond)

    @classmethod
    def random(cls, seed: int, device, cond: bool = False) -> "NerfNetwork":
        return cls.from_state_dict(init_state_dict(seed, cond), device, cond)

    def repack(self) -> None:
        """Refresh the packed blob after ``flat`` changed (optimizer step)."""
        ops.pack_weights(self.flat, self.cond, out=self.packed)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: v.clone() for k, v in unflatten(self.flat.detach(), self.cond).items()}


class NerfModel:
    """Coarse + fine networks (A.9)."""

    def __init__(self, coarse: NerfNetwork, fine: Optional[NerfNetwork]):
        self.coarse, self.fine = coarse, fine if fine is not None else coarse
        self.cond = coarse.cond

    @classmethod
    def random(cls, device, cond: bool = False, seeds=(0, 1)) -> "NerfModel":
        return cls(NerfNetwork.random(seeds[0], device, cond), NerfNetwork.random(seeds[1], device, cond))

    @property
    def device(self):
        return self.coarse.flat.device

    def repack(self):
        self.coarse.repack()
        if self.fine is not self.coarse:
            self.fine.repack()
