"""Stage-level operators over the C ABI (include/fnerf.h).  Every function takes CUDA fp32 tensors,
allocates its outputs with torch (the library never allocates) and enqueues on torch's current
stream.  CPU tensors are rejected: there is no fallback path."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import PRECISION_BF16, PRECISION_FP32, check

PRECISIONS = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.FnerfError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def param_count(cond: bool = False) -> int:
    return int(_lib.load().fnerf_param_count(int(cond)))


def packed_bytes(cond: bool = False) -> int:
    return int(_lib.load().fnerf_packed_bytes(int(cond)))


def pack_weights(flat: torch.Tensor, cond: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """flat fp32 [param_count] (nn.Linear order, include/fnerf.h) -> packed uint8 blob."""
    flat = _f32(flat, "flat")
    assert flat.numel() == param_count(cond), (flat.numel(), param_count(cond))
    if out is None:
        out = torch.empty(packed_bytes(cond), dtype=torch.uint8, device=flat.device)
    with torch.cuda.device(flat.device):
        check(_lib.load().fnerf_pack_weights(flat.data_ptr(), out.data_ptr(), int(cond), _stream()), "pack_weights")
    return out


def unpack_weights(packed: torch.Tensor, cond: bool = False) -> torch.Tensor:
    flat = torch.empty(param_count(cond), dtype=torch.float32, device=packed.device)
    with torch.cuda.device(packed.device):
        check(_lib.load().fnerf_unpack_weights(packed.data_ptr(), flat.data_ptr(), int(cond), _stream()), "unpack_weights")
    return flat


def ray_setup(rays_d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    rays_d = _f32(rays_d, "rays_d")
    R = rays_d.shape[0]
    vd = torch.empty_like(rays_d)
    dn = torch.empty(R, dtype=torch.float32, device=rays_d.device)
    with torch.cuda.device(rays_d.device):
        check(_lib.load().fnerf_ray_setup(rays_d.data_ptr(), vd.data_ptr(), dn.data_ptr(), R, _stream()), "ray_setup")
    return vd, dn


def stratified(near: torch.Tensor, far: torch.Tensor, t_vals: torch.Tensor,
               u_strat: Optional[torch.Tensor] = None, lindisp: bool = False) -> torch.Tensor:
    near, far, t_vals = _f32(near, "near"), _f32(far, "far"), _f32(t_vals, "t_vals")
    R, N = near.numel(), t_vals.numel()
    if u_strat is not None:
        u_strat = _f32(u_strat, "u_strat")
        assert u_strat.shape == (R, N)
    z = torch.empty(R, N, dtype=torch.float32, device=near.device)
    with torch.cuda.device(near.device):
        check(_lib.load().fnerf_stratified(near.data_ptr(), far.data_ptr(), t_vals.data_ptr(), _ptr(u_strat),
                                           z.data_ptr(), R, N, int(lindisp), _stream()), "stratified")
    return z


def importance(z_c: torch.Tensor, weights_c: torch.Tensor, u: torch.Tensor, want_idx: bool = True):
    """A.7.  u: [R,Nf] or [Nf] (shared row).  Returns dict(z_samples, z_f, inds, z_std)."""
    z_c, weights_c, u = _f32(z_c, "z_c"), _f32(weights_c, "weights_c"), _f32(u, "u")
    R, Nc = z_c.shape
    Nf = u.shape[-1]
    stride = 0 if u.dim() == 1 else Nf
    dev = z_c.device
    z_samples = torch.empty(R, Nf, dtype=torch.float32, device=dev)
    z_f = torch.empty(R, Nc + Nf, dtype=torch.float32, device=dev)
    inds = torch.empty(R, Nf, dtype=torch.int32, device=dev) if want_idx else None
    z_std = torch.empty(R, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().fnerf_importance(z_c.data_ptr(), weights_c.data_ptr(), u.data_ptr(), stride,
                                           z_samples.data_ptr(), z_f.data_ptr(), _ptr(inds), z_std.data_ptr(),
                                           R, Nc, Nf, _stream()), "importance")
    return {"z_samples": z_samples, "z_f": z_f, "inds": inds, "z_std": z_std}


def posenc(x: torch.Tensor, L: int) -> torch.Tensor:
    x = _f32(x, "x")
    M = x.shape[0]
    out = torch.empty(M, 3 + 6 * L, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().fnerf_posenc(x.data_ptr(), out.data_ptr(), M, L, _stream()), "posenc")
    return out


def cond_project(packed: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    cond = _f32(cond, "cond").reshape(-1, 256)
    proj = torch.empty_like(cond)
    with torch.cuda.device(cond.device):
        check(_lib.load().fnerf_cond_project(packed.data_ptr(), cond.data_ptr(), proj.data_ptr(), cond.shape[0],
                                             _stream()), "cond_project")
    return proj


def mlp_fwd(packed: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, viewdirs: torch.Tensor,
            z: torch.Tensor, *, precision: str = "bf16", cond_proj: Optional[torch.Tensor] = None,
            cond_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused network query: raw[R,S,4] for pts = rays_o + rays_d * z (A.3+A.4+A.8)."""
    rays_o, rays_d, viewdirs, z = (_f32(rays_o, "rays_o"), _f32(rays_d, "rays_d"), _f32(viewdirs, "viewdirs"),
                                   _f32(z, "z"))
    R, S = z.shape
    raw = torch.empty(R, S, 4, dtype=torch.float32, device=z.device)
    has_cond = cond_proj is not None
    C = cond_proj.shape[0] if has_cond else 0
    if cond_index is not None:
        cond_index = cond_index.to(torch.int32).contiguous()
    with torch.cuda.device(z.device):
        check(_lib.load().fnerf_mlp_fwd(PRECISIONS[precision], packed.data_ptr(), int(has_cond), rays_o.data_ptr(),
                                        rays_d.data_ptr(), viewdirs.data_ptr(), z.data_ptr(), _ptr(cond_proj),
                                        _ptr(cond_index), C, raw.data_ptr(), R, S, _stream()), "mlp_fwd")
    return raw


def mlp_fwd_composite_supported(S: int) -> bool:
    """True when the fused query + compositing kernel serves S samples per ray (whole-ray tile groups, include/fnerf.h)."""
    return bool(_lib.load().fnerf_mlp_fwd_composite_supported(int(S)))


def mlp_fwd_composite(packed: torch.Tensor, rays_o, rays_d, viewdirs, dnorm, z, *, cond_proj=None, cond_index=None,
                      raw_noise: Optional[torch.Tensor] = None, white_bkgd: bool = False, want_raw: bool = False,
                      want_weights: bool = True):
    """bf16 network query with A.5 fused into its last epilogue: dict(rgb, depth, acc, disp, weights, raw);
    raw[R,S,4] is only produced (and only touches HBM) when want_raw."""
    rays_o, rays_d, viewdirs, dnorm, z = (_f32(rays_o, "rays_o"), _f32(rays_d, "rays_d"), _f32(viewdirs, "viewdirs"),
                                          _f32(dnorm, "dnorm"), _f32(z, "z"))
    R, S = z.shape
    dev = z.device
    if raw_noise is not None:
        raw_noise = _f32(raw_noise, "raw_noise")
        assert raw_noise.shape == (R, S)
    has_cond = cond_proj is not None
    C = cond_proj.shape[0] if has_cond else 0
    if cond_index is not None:
        cond_index = cond_index.to(torch.int32).contiguous()

    def new(*shape):
        return torch.empty(*shape, dtype=torch.float32, device=dev)

    raw = new(R, S, 4) if want_raw else None
    weights = new(R, S) if want_weights else None
    rgb, depth, acc, disp = new(R, 3), new(R), new(R), new(R)
    with torch.cuda.device(dev):
        check(_lib.load().fnerf_mlp_fwd_composite(packed.data_ptr(), int(has_cond), rays_o.data_ptr(), rays_d.data_ptr(),
                                                  viewdirs.data_ptr(), dnorm.data_ptr(), z.data_ptr(), _ptr(cond_proj),
                                                  _ptr(cond_index), C, _ptr(raw_noise), _ptr(raw), rgb.data_ptr(),
                                                  depth.data_ptr(), acc.data_ptr(), disp.data_ptr(), _ptr(weights), R, S,
                                                  int(white_bkgd), _stream()), "mlp_fwd_composite")
    return {"rgb": rgb, "depth": depth, "acc": acc, "disp": disp, "weights": weights, "raw": raw}


def mlp_bwd(packed: torch.Tensor, rays_o, rays_d, viewdirs, z, g_raw: torch.Tensor, flat_grad: torch.Tensor, *,
            precision: str = "fp32", cond_rows=None, cond_index=None) -> torch.Tensor:
    """Accumulates dL/dparams into flat_grad (flat layout) given g_raw[R,S,4].  cond_rows: RAW codes [C,256]."""
    rays_o, rays_d, viewdirs, z, g_raw = (_f32(rays_o, "rays_o"), _f32(rays_d, "rays_d"),
                                          _f32(viewdirs, "viewdirs"), _f32(z, "z"), _f32(g_raw, "g_raw"))
    R, S = z.shape
    lib = _lib.load()
    has_cond = cond_rows is not None
    ws_bytes = int(lib.fnerf_mlp_bwd_workspace_bytes(PRECISIONS[precision], int(has_cond), R, S))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=z.device)
    if has_cond:
        cond_rows = _f32(cond_rows, "cond_rows").reshape(-1, 256)
    C = cond_rows.shape[0] if has_cond else 0
    if cond_index is not None:
        cond_index = cond_index.to(torch.int32).contiguous()
    with torch.cuda.device(z.device):
        check(lib.fnerf_mlp_bwd(PRECISIONS[precision], packed.data_ptr(), int(has_cond), rays_o.data_ptr(),
                                rays_d.data_ptr(), viewdirs.data_ptr(), z.data_ptr(), _ptr(cond_rows),
                                _ptr(cond_index), C, g_raw.data_ptr(), flat_grad.data_ptr(), ws.data_ptr(), ws_bytes,
                                R, S, _stream()), "mlp_bwd")
    return flat_grad


def adam_step(params: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, *,
              lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    """Fused in-place Adam step over flat fp32 CUDA buffers (A.10)."""
    for t in (params, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.FnerfError("adam_step needs contiguous CUDA float32 buffers")
    with torch.cuda.device(params.device):
        check(_lib.load().fnerf_adam_step(params.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                          params.numel(), lr, betas[0], betas[1], eps, step, grad_scale, _stream()), "adam_step")


def allreduce_adam_step(peer_grads_dev: int, world: int, offset: int, params: torch.Tensor, exp_avg: torch.Tensor,
                        exp_avg_sq: torch.Tensor, step: int, *, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """Fused mean-all-reduce + Adam over peer memory.  peer_grads_dev: device address of an array of `world` pointers to
    the ranks' flat gradient buffers (e.g. torch symmetric memory's ``buffer_ptrs_dev``); elements
    [offset, offset + params.numel()) of every buffer are summed in rank order."""
    for t in (params, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.FnerfError("allreduce_adam_step needs contiguous CUDA float32 buffers")
    with torch.cuda.device(params.device):
        check(_lib.load().fnerf_allreduce_adam_step(peer_grads_dev, world, offset, params.data_ptr(), exp_avg.data_ptr(),
                                                    exp_avg_sq.data_ptr(), params.numel(), lr, betas[0], betas[1], eps, step,
                                                    1.0 / world, _stream()), "allreduce_adam_step")


def multimem_allreduce(multicast_ptr: int, rank: int, world: int, n: int, device) -> None:
    """In-place sum over the ranks' buffers behind an NVSwitch multicast mapping (NVLS); n % 4 == 0."""
    with torch.cuda.device(device):
        check(_lib.load().fnerf_multimem_allreduce(multicast_ptr, rank, world, n, _stream()), "multimem_allreduce")


def mlp_tape_bytes(R: int, S: int) -> int:
    return int(_lib.load().fnerf_mlp_tape_bytes(R, S))


def mlp_fwd_tape(packed: torch.Tensor, rays_o, rays_d, viewdirs, z, *, cond_proj=None, cond_index=None):
    """Training forward (bf16): returns (raw[R,S,4], tape) -- the tape is an opaque uint8 tensor for mlp_bwd_tape."""
    rays_o, rays_d, viewdirs, z = (_f32(rays_o, "rays_o"), _f32(rays_d, "rays_d"), _f32(viewdirs, "viewdirs"),
                                   _f32(z, "z"))
    R, S = z.shape
    raw = torch.empty(R, S, 4, dtype=torch.float32, device=z.device)
    tape = torch.empty(max(mlp_tape_bytes(R, S), 16), dtype=torch.uint8, device=z.device)
    has_cond = cond_proj is not None
    C = cond_proj.shape[0] if has_cond else 0
    if cond_index is not None:
        cond_index = cond_index.to(torch.int32).contiguous()
    with torch.cuda.device(z.device):
        check(_lib.load().fnerf_mlp_fwd_tape(packed.data_ptr(), int(has_cond), rays_o.data_ptr(), rays_d.data_ptr(),
                                             viewdirs.data_ptr(), z.data_ptr(), _ptr(cond_proj), _ptr(cond_index), C,
                                             raw.data_ptr(), tape.data_ptr(), tape.numel(), R, S, _stream()), "mlp_fwd_tape")
    return raw, tape


def mlp_bwd_tape(packed: torch.Tensor, g_raw: torch.Tensor, tape: torch.Tensor, flat_grad: torch.Tensor, *,
                 cond_rows=None, cond_index=None) -> torch.Tensor:
    """flat_grad += dL/dparams from the tape of mlp_fwd_tape / render_rays(save_tape=True) and g_raw[R,S,4].
    cond_rows: RAW codes [C,256] of a conditioned network."""
    g_raw = _f32(g_raw, "g_raw")
    R, S = g_raw.shape[:2]
    has_cond = cond_rows is not None
    if has_cond:
        cond_rows = _f32(cond_rows, "cond_rows").reshape(-1, 256)
    C = cond_rows.shape[0] if has_cond else 0
    if cond_index is not None:
        cond_index = cond_index.to(torch.int32).contiguous()
    lib = _lib.load()
    ws_bytes = int(lib.fnerf_mlp_bwd_tape_workspace_bytes(R, S))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=g_raw.device)
    with torch.cuda.device(g_raw.device):
        check(lib.fnerf_mlp_bwd_tape(packed.data_ptr(), int(has_cond), g_raw.data_ptr(), tape.data_ptr(), tape.numel(),
                                     _ptr(cond_rows), _ptr(cond_index), C, flat_grad.data_ptr(), ws.data_ptr(), ws.numel(),
                                     R, S, _stream()), "mlp_bwd_tape")
    return flat_grad


def composite_fwd(raw: torch.Tensor, z: torch.Tensor, dnorm: torch.Tensor, *, white_bkgd: bool = False,
                  raw_noise: Optional[torch.Tensor] = None, want_weights: bool = True):
    """A.5 raw2outputs -> dict(rgb, depth, acc, disp, weights)."""
    raw, z, dnorm = _f32(raw, "raw"), _f32(z, "z"), _f32(dnorm, "dnorm")
    R, S = z.shape
    dev = z.device
    if raw_noise is not None:
        raw_noise = _f32(raw_noise, "raw_noise")
    rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
    depth = torch.empty(R, dtype=torch.float32, device=dev)
    acc = torch.empty(R, dtype=torch.float32, device=dev)
    disp = torch.empty(R, dtype=torch.float32, device=dev)
    weights = torch.empty(R, S, dtype=torch.float32, device=dev) if want_weights else None
    with torch.cuda.device(dev):
        check(_lib.load().fnerf_composite_fwd(raw.data_ptr(), z.data_ptr(), dnorm.data_ptr(), _ptr(raw_noise),
                                              rgb.data_ptr(), depth.data_ptr(), acc.data_ptr(), disp.data_ptr(),
                                              _ptr(weights), R, S, int(white_bkgd), _stream()), "composite_fwd")
    return {"rgb": rgb, "depth": depth, "acc": acc, "disp": disp, "weights": weights}


def composite_bwd(raw: torch.Tensor, z: torch.Tensor, dnorm: torch.Tensor, g_rgb: torch.Tensor,
                  g_depth: Optional[torch.Tensor] = None, g_acc: Optional[torch.Tensor] = None, *,
                  white_bkgd: bool = False, raw_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """A.6 -> g_raw[R,S,4].  raw_noise: the tensor composite_fwd was given (sigma = raw[...,3] + raw_noise)."""
    raw, z, dnorm, g_rgb = _f32(raw, "raw"), _f32(z, "z"), _f32(dnorm, "dnorm"), _f32(g_rgb, "g_rgb")
    R, S = z.shape
    if g_depth is not None:
        g_depth = _f32(g_depth, "g_depth")
    if g_acc is not None:
        g_acc = _f32(g_acc, "g_acc")
    if raw_noise is not None:
        raw_noise = _f32(raw_noise, "raw_noise")
        assert raw_noise.shape == (R, S)
    g_raw = torch.empty(R, S, 4, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        check(_lib.load().fnerf_composite_bwd(raw.data_ptr(), z.data_ptr(), dnorm.data_ptr(), _ptr(raw_noise), g_rgb.data_ptr(),
                                              _ptr(g_depth), _ptr(g_acc), g_raw.data_ptr(), R, S, int(white_bkgd),
                                              _stream()), "composite_bwd")
    return g_raw
