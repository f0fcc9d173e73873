"""Data-parallel training step: render_rays forward, loss = mse(rgb, tgt) +
mse(rgb0, tgt), backward through compositing (A.6) and the MLP, ONE all-reduce over the flat fp32
gradient buffer of both networks (1,191,688 floats = 4.77 MB), Adam on the fp32 master parameters,
re-pack of the kernel blobs.  One process per GPU; ``torch.distributed`` (NCCL over NVLink on the
B200 box, gloo in the CPU tests) is used for nothing but that all-reduce.

The host-side control flow (`allreduce_mean_`, `FlatAdam`'s step bookkeeping) is plain torch so the world_size>1 logic
is testable on CPU with gloo; every arithmetic step -- gradients AND the optimiser update -- is a CUDA kernel of
libfnerf.so.  (The CPU tests inject their own update function into `FlatAdam`; the product has no CPU mirror.)"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.distributed as dist

from .model import NerfModel
from .render import render_rays


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """In-place average of the flat gradient buffer over the process group (no-op when not initialised)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
            flat_grad.div_(world)
    return flat_grad


def _fused_adam_kernel(params, grad, m, v, t, lr, b1, b2, eps, grad_scale):
    from . import ops
    ops.adam_step(params, grad.contiguous(), m, v, t, lr=lr, betas=(b1, b2), eps=eps, grad_scale=grad_scale)


class FlatAdam:
    """Adam (torch defaults: betas 0.9/0.999, eps 1e-8; A.10) over one flat fp32 parameter buffer.  The update itself
    is `update_fn(params, grad, m, v, t, lr, b1, b2, eps, grad_scale)`: the fused CUDA kernel ``fnerf_adam_step`` unless
    a test injects another one."""

    def __init__(self, n: int, device, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, update_fn=None):
        self.update_fn = update_fn or _fused_adam_kernel
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.m = torch.zeros(n, dtype=torch.float32, device=device)
        self.v = torch.zeros(n, dtype=torch.float32, device=device)
        self.t = 0

    @torch.no_grad()
    def step(self, params: torch.Tensor, grad: torch.Tensor) -> None:
        """One Adam step over the whole buffer."""
        self.t += 1
        self.apply(params, grad, 0)

    @torch.no_grad()
    def apply(self, params: torch.Tensor, grad: torch.Tensor, offset: int, grad_scale: float = 1.0) -> None:
        """Update `params` (a view of the flat buffer starting at `offset`) at the current step count."""
        n = params.numel()
        m, v = self.m[offset:offset + n], self.v[offset:offset + n]
        self.update_fn(params, grad, m, v, self.t, self.lr, self.b1, self.b2, self.eps, grad_scale)


class Trainer:
    """Holds the optimizer state of a NerfModel and runs data-parallel steps."""

    def __init__(self, model: NerfModel, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, group=None,
                 fused_allreduce=None):
        """fused_allreduce: sum the ranks' gradients with P2P loads over NVLink inside the Adam kernel
        (``fnerf_allreduce_adam_step`` on a torch symmetric-memory gradient buffer) instead of NCCL all-reduce + Adam.
        None = use it for 2..4 ranks when the symmetric-memory rendezvous succeeds: the one-shot kernel reads every
        peer's whole buffer (measured: 32 us vs 50 us for NCCL + Adam at 2 GPUs, 74 us vs 67 us at 8, where NCCL's
        NVLS / ring schedule moves less data per rank).  "nvls": sum in the NVSwitch instead
        (``fnerf_multimem_allreduce`` on the buffer's multicast mapping, then the plain fused Adam); bit-identical to NCCL's
        result in the 2- and 8-GPU checks, 60 / 75 us."""
        self.model, self.group = model, group
        self.n_c, self.n_f = model.coarse.flat.numel(), model.fine.flat.numel()
        self.shared = model.fine is model.coarse
        n = self.n_c if self.shared else self.n_c + self.n_f
        self.opt = FlatAdam(n, model.device, lr, betas, eps)
        self.symm = None
        self.nvls = False
        self.flat_grad = None
        self.events = None                        # profiling hook of reduce_and_update()
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world = world
        if world > 1 and (fused_allreduce or (fused_allreduce is None and world <= 4)) and model.device.type == "cuda":
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                buf = symm_mem.empty(n, dtype=torch.float32, device=model.device)
                self.symm = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                self.nvls = fused_allreduce == "nvls"
                if self.nvls and (not self.symm.multicast_ptr or n % 4):
                    raise RuntimeError("no NVSwitch multicast mapping for the gradient buffer")
            except Exception as e:                     # no P2P / fabric handles on this rank
                err = e
            # the ranks must AGREE on the mode: a rank that fell back to NCCL while its peers wait in a symmetric-memory
            # barrier would deadlock the job
            ok = torch.tensor([0 if err is not None else 1], device=model.device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 1:
                self.flat_grad = buf.zero_()
            else:
                self.symm, self.nvls = None, False
                if fused_allreduce:
                    raise RuntimeError(f"fused all-reduce ({fused_allreduce!r}) unavailable on at least one rank"
                                       + (f": {err}" if err is not None else ""))
        if self.flat_grad is None:
            self.flat_grad = torch.zeros(n, dtype=torch.float32, device=model.device)   # the all-reduce buffer

    def step(self, rays_o: torch.Tensor, rays_d: torch.Tensor, target: torch.Tensor, near, far, N_samples: int,
             N_importance: int, cond: Optional[torch.Tensor] = None, *, view_id=None, u_strat=None, u_fine=None,
             white_bkgd: bool = False, precision: str = "bf16") -> Dict[str, torch.Tensor]:
        m = self.model
        fc = m.coarse.flat.detach().requires_grad_(True)
        ff = fc if self.shared else m.fine.flat.detach().requires_grad_(True)
        m.coarse.flat, m.fine.flat = fc, ff
        try:
            out = render_rays(m, rays_o, rays_d, near, far, N_samples, N_importance, cond, view_id=view_id,
                              u_strat=u_strat, u_fine=u_fine, white_bkgd=white_bkgd, precision=precision)
            mse = torch.nn.functional.mse_loss                   # one fused kernel each way instead of sub / pow / mean / ...
            loss_f = mse(out["rgb"], target)
            loss = loss_f + mse(out["rgb0"], target) if N_importance > 0 else loss_f
            loss.backward()
        finally:                                  # never leave autograd leaves behind in the model
            m.coarse.flat, m.fine.flat = fc.detach(), ff.detach()
        with torch.no_grad():
            g = self.flat_grad
            g[:self.n_c].copy_(fc.grad)
            if not self.shared:
                if ff.grad is not None:
                    g[self.n_c:].copy_(ff.grad)
                else:
                    g[self.n_c:].zero_()
        self.reduce_and_update()
        return {"loss": loss.detach(), "psnr": -10.0 * torch.log10(loss_f.detach())}

    @torch.no_grad()
    def reduce_and_update(self) -> None:
        """The data-parallel half of a step, on whatever local gradients `self.flat_grad` holds: mean over the ranks (the
        ONE collective of the system), Adam on the fp32 master parameters, re-pack of the kernel blobs.
        `self.events`, when set to a list, receives one (start, stop) CUDA-event pair around reduce + Adam per call."""
        m, g, o = self.model, self.flat_grad, self.opt
        ev = None
        if self.events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        o.t += 1
        if self.symm is not None:
            # every rank's gradients are in its symmetric buffer: barrier, then each rank sums all peers' buffers over
            # NVLink inside the Adam kernel (same rank order everywhere -> identical replicas), barrier again so nobody
            # overwrites its buffer while a peer still reads it
            from . import ops
            self.symm.barrier(channel=0)
            if self.nvls:
                ops.multimem_allreduce(self.symm.multicast_ptr, self.symm.rank, self.symm.world_size, g.numel(), g.device)
                self.symm.barrier(channel=1)
                sc = 1.0 / self.symm.world_size
                o.apply(m.coarse.flat, g[:self.n_c], 0, sc)
                if not self.shared:
                    o.apply(m.fine.flat, g[self.n_c:], self.n_c, sc)
            else:
                ops.allreduce_adam_step(self.symm.buffer_ptrs_dev, self.symm.world_size, 0, m.coarse.flat, o.m[:self.n_c],
                                        o.v[:self.n_c], o.t, lr=o.lr, betas=(o.b1, o.b2), eps=o.eps)
                if not self.shared:
                    ops.allreduce_adam_step(self.symm.buffer_ptrs_dev, self.symm.world_size, self.n_c, m.fine.flat,
                                            o.m[self.n_c:], o.v[self.n_c:], o.t, lr=o.lr, betas=(o.b1, o.b2), eps=o.eps)
                self.symm.barrier(channel=1)
        else:
            allreduce_mean_(g, self.group)
            o.apply(m.coarse.flat, g[:self.n_c], 0)
            if not self.shared:
                o.apply(m.fine.flat, g[self.n_c:], self.n_c)
        if ev is not None:
            ev[1].record()
            self.events.append(ev)
        m.repack()

    @property
    def mode(self) -> str:
        """How this trainer sums gradients over the ranks: 'single' | 'nccl' | 'p2p' (fused one-shot kernel) | 'nvls'."""
        if self.world <= 1:
            return "single"
        return "nccl" if self.symm is None else ("nvls" if self.nvls else "p2p")


def psnr_from_mse(mse: float) -> float:
    return -10.0 * math.log10(max(mse, 1e-20))
