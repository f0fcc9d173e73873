"""render_rays: the Python operator boundary.

``render_rays(model, rays_o, rays_d, near, far, N_samples, N_importance, cond=None, ...)`` is the
public call; it lowers to the torch custom op ``fnerf::render_rays`` (fake-tensor and autograd
registered), which makes ONE C-ABI call, ``fnerf_render_rays``, that enqueues the whole
coarse -> composite -> importance -> fine -> composite chain on torch's current stream.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib, ops
from .model import NerfModel

_T = torch.Tensor
_tvals_cache: Dict[tuple, torch.Tensor] = {}
_profile_events = None      # bench.py's profiling hook (set_profile_events); not part of the render contract


def set_profile_events(events) -> None:
    """Install (or clear with None) four already-recorded torch.cuda.Event(enable_timing=True) objects;
    every following render_rays call records them around its coarse / fine network-query launches."""
    global _profile_events
    _profile_events = None if events is None else tuple(events)



def _linspace01(n: int, device) -> torch.Tensor:
    """torch.linspace evaluated on the CPU (the oracle's bits: linspace(0,1,n)[i] != i/(n-1)) and
    cached on the device."""
    key = (n, str(device))
    t = _tvals_cache.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, n).to(device)
        _tvals_cache[key] = t
    return t


def _render_impl(packed_c, packed_f, rays_o, rays_d, near, far, t_vals, u_strat, u_fine, noise_c, noise_f, cond_proj_c,
                 cond_proj_f, cond_index, n_importance, white_bkgd, lindisp, precision, save_tape=False, fuse=False, taps=True) -> List[_T]:
    lib = _lib.load()
    dev = rays_o.device
    R, Nc, Nf = rays_o.shape[0], t_vals.numel(), int(n_importance)
    S = Nc + Nf

    def new(*shape):
        return torch.empty(*shape, dtype=torch.float32, device=dev)

    rgb, disp, acc, depth = new(R, 3), new(R), new(R), new(R)
    rgb0, disp0, acc0, z_std, depth0 = new(R, 3), new(R), new(R), new(R), new(R)
    # fused compositing without taps: raw[R,S,4] is neither allocated nor written
    want_raw = taps or not fuse
    z_c, raw_c, w_c = new(R, Nc), (new(R, Nc, 4) if want_raw else new(0)), new(R, Nc)
    z_f, raw_f = (new(R, S), (new(R, S, 4) if want_raw else new(0))) if Nf > 0 else (new(0), new(0))
    ws_bytes = int(lib.fnerf_render_rays_workspace_bytes(R, Nc, Nf))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)

    a = _lib.RenderArgs()
    p = ops._ptr
    a.packed_coarse, a.packed_fine = packed_c.data_ptr(), packed_f.data_ptr()
    a.cond = int(cond_proj_c is not None)
    a.precision = int(precision)
    a.rays_o, a.rays_d, a.near, a.far = rays_o.data_ptr(), rays_d.data_ptr(), near.data_ptr(), far.data_ptr()
    a.t_vals, a.u_strat = t_vals.data_ptr(), p(u_strat)
    a.u_fine = p(u_fine)
    a.u_fine_row_stride = 0 if (u_fine is None or u_fine.dim() == 1) else Nf
    a.raw_noise_coarse, a.raw_noise_fine = p(noise_c), p(noise_f)
    a.cond_proj_coarse, a.cond_proj_fine, a.cond_index = p(cond_proj_c), p(cond_proj_f), p(cond_index)
    a.C = 0 if cond_proj_c is None else cond_proj_c.shape[0]
    a.R, a.Nc, a.Nf = R, Nc, Nf
    a.white_bkgd, a.lindisp = int(white_bkgd), int(lindisp)
    a.rgb, a.disp, a.acc, a.depth = rgb.data_ptr(), disp.data_ptr(), acc.data_ptr(), depth.data_ptr()
    a.rgb0, a.disp0, a.acc0, a.z_std = rgb0.data_ptr(), disp0.data_ptr(), acc0.data_ptr(), z_std.data_ptr()
    a.depth0 = depth0.data_ptr()
    a.z_c, a.weights_c = z_c.data_ptr(), w_c.data_ptr()
    a.raw_c = raw_c.data_ptr() if want_raw else None
    if Nf > 0:
        a.z_f = z_f.data_ptr()
        a.raw_f = raw_f.data_ptr() if want_raw else None
    a.weights_f = None
    a.fuse_composite = int(fuse)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    # training forward: activations taped for the tensor-core backward.  The tapes are op OUTPUTS (uint8, hence
    # non-differentiable: autograd never materialises gradients for them) that only the autograd node keeps alive.
    tape_c, tape_f = torch.empty(0, dtype=torch.uint8, device=dev), torch.empty(0, dtype=torch.uint8, device=dev)
    if save_tape:
        tape_c = torch.empty(max(ops.mlp_tape_bytes(R, Nc), 16), dtype=torch.uint8, device=dev)
        a.tape_coarse, a.tape_coarse_bytes = tape_c.data_ptr(), tape_c.numel()
        if Nf > 0:
            tape_f = torch.empty(max(ops.mlp_tape_bytes(R, S), 16), dtype=torch.uint8, device=dev)
            a.tape_fine, a.tape_fine_bytes = tape_f.data_ptr(), tape_f.numel()
    if _profile_events is not None:      # (coarse_start, coarse_stop, fine_start, fine_stop) torch.cuda.Event
        a.ev_coarse_start, a.ev_coarse_stop, a.ev_fine_start, a.ev_fine_stop = [e.cuda_event for e in _profile_events]
    with torch.cuda.device(dev):
        _lib.check(lib.fnerf_render_rays(ctypes.byref(a), torch.cuda.current_stream().cuda_stream), "render_rays")
    return [rgb, disp, acc, depth, rgb0, disp0, acc0, z_std, z_c, z_f, raw_c, raw_f, depth0, tape_c, tape_f]


@torch.library.custom_op("fnerf::render_rays", mutates_args=())
def render_rays_op(flat_c: _T, flat_f: _T, packed_c: _T, packed_f: _T, rays_o: _T, rays_d: _T, near: _T, far: _T,
                   t_vals: _T, u_strat: Optional[_T], u_fine: Optional[_T], noise_c: Optional[_T], noise_f: Optional[_T],
                   cond_proj_c: Optional[_T], cond_proj_f: Optional[_T], cond_index: Optional[_T], cond_rows: Optional[_T],
                   n_importance: int, white_bkgd: bool, lindisp: bool, precision: int, save_tape: bool, fuse: bool,
                   taps: bool) -> List[_T]:
    # flat_c / flat_f only anchor the autograd graph (and cond_rows only feeds the backward); the
    # forward kernels read the packed blobs and the hoisted projections.
    return _render_impl(packed_c, packed_f, rays_o, rays_d, near, far, t_vals, u_strat, u_fine, noise_c, noise_f, cond_proj_c,
                        cond_proj_f, cond_index, n_importance, white_bkgd, lindisp, precision, save_tape, fuse, taps)


@render_rays_op.register_fake
def _(flat_c, flat_f, packed_c, packed_f, rays_o, rays_d, near, far, t_vals, u_strat, u_fine, noise_c, noise_f, cond_proj_c,
      cond_proj_f, cond_index, cond_rows, n_importance, white_bkgd, lindisp, precision, save_tape, fuse, taps):
    R, Nc, S = rays_o.shape[0], t_vals.numel(), t_vals.numel() + n_importance
    e = rays_o.new_empty
    want_raw = taps or not fuse
    zf, rf = (e(R, S), (e(R, S, 4) if want_raw else e(0))) if n_importance > 0 else (e(0), e(0))
    d0 = e(R)

    def tape(samples):
        return rays_o.new_empty(max(ops.mlp_tape_bytes(R, samples), 16) if (save_tape and samples) else 0, dtype=torch.uint8)

    return [e(R, 3), e(R), e(R), e(R), e(R, 3), e(R), e(R), e(R), e(R, Nc), zf, (e(R, Nc, 4) if want_raw else e(0)), rf, d0,
            tape(Nc), tape(S if n_importance > 0 else 0)]


_N_INPUTS = 24


def _setup_context(ctx, inputs, output):
    (flat_c, flat_f, packed_c, packed_f, rays_o, rays_d, near, far, t_vals, u_strat, u_fine, noise_c, noise_f, cond_proj_c,
     cond_proj_f, cond_index, cond_rows, n_importance, white_bkgd, lindisp, precision, save_tape, fuse, taps) = inputs
    ctx.set_materialize_grads(False)          # unused outputs arrive as None: their backward branch is skipped
    rgb, disp, acc, depth, rgb0, disp0, acc0 = output[:7]
    z_c, z_f, raw_c, raw_f, depth0, tape_c, tape_f = output[8:15]
    ctx.save_for_backward(packed_c, packed_f, rays_o, rays_d, z_c, z_f, raw_c, raw_f, cond_rows, cond_index, noise_c, noise_f,
                          disp, acc, depth, disp0, acc0, depth0, tape_c, tape_f)
    ctx.n_importance, ctx.white_bkgd, ctx.precision, ctx.save_tape = n_importance, white_bkgd, precision, save_tape
    ctx.n_c, ctx.n_f = flat_c.numel(), flat_f.numel()
    ctx.has_raw = taps or not fuse


def _disp_chain(g_disp, disp, acc, depth, g_depth, g_acc):
    """disp = 1 / max(1e-10, depth / acc)  ->  adds dL/ddisp to the depth / acc gradients (zero where the clamp is active)."""
    q = depth / acc
    live = (q > 1e-10).to(disp.dtype)
    gq = -g_disp * disp * disp * live
    gd = gq / acc
    ga = -gq * q / acc
    gd, ga = torch.nan_to_num(gd, nan=0.0, posinf=0.0, neginf=0.0), torch.nan_to_num(ga, nan=0.0, posinf=0.0, neginf=0.0)
    return (gd if g_depth is None else g_depth + gd), (ga if g_acc is None else g_acc + ga)


def _backward(ctx, grads):
    """A.6 + MLP backward.  Gradients are accepted for the maps rgb / disp / acc / depth (fine) and rgb0 / disp0 / acc0
    (coarse) and for the raw_c / raw_f taps; sample positions are detached (A.7), so z_std and the z taps carry none and a
    gradient arriving for them is an error rather than a silent zero."""
    (packed_c, packed_f, rays_o, rays_d, z_c, z_f, raw_c, raw_f, cond_rows, cidx, noise_c, noise_f,
     disp, acc, depth, disp0, acc0, depth0, tape_c, tape_f) = ctx.saved_tensors
    g_rgb, g_disp, g_acc, g_depth, g_rgb0, g_disp0, g_acc0, g_zstd, g_zc, g_zf, g_rawc, g_rawf = grads[:12]
    if not ctx.has_raw:
        raise RuntimeError("render_rays: this forward ran with compositing fused into the network query and kept no raw "
                           "tensor (fuse_composite=True, return_taps=False); A.6 needs it -- render with fuse_composite=False")
    if g_zstd is not None or g_zc is not None or g_zf is not None:
        raise RuntimeError("render_rays: z_std / z_c / z_f are detached sample positions (SURVEY.md A.7) and carry no gradient")
    viewdirs, dnorm = ops.ray_setup(rays_d)
    R = rays_o.shape[0]
    dev = rays_o.device

    def grad_net(packed, z, raw, noise, gr, gd, ga, g_tap, n_params, tape):
        if gr is None and gd is None and ga is None and g_tap is None:
            return None
        g_raw = None
        if not (gr is None and gd is None and ga is None):
            if gr is None:
                gr = torch.zeros(R, 3, device=dev)
            g_raw = ops.composite_bwd(raw, z, dnorm, gr.contiguous(), gd, ga, white_bkgd=ctx.white_bkgd, raw_noise=noise)
        if g_tap is not None:
            g_raw = g_tap.contiguous() if g_raw is None else g_raw + g_tap
        flat_grad = torch.zeros(n_params, dtype=torch.float32, device=dev)
        if ctx.save_tape:        # the forward taped its activations: dgrad + wgrad straight from the tape
            ops.mlp_bwd_tape(packed, g_raw, tape, flat_grad, cond_rows=cond_rows, cond_index=cidx)
            return flat_grad
        # otherwise recompute, with the same arithmetic as the forward that produced `raw`
        ops.mlp_bwd(packed, rays_o, rays_d, viewdirs, z, g_raw, flat_grad, precision=_PRECISION_NAMES[ctx.precision],
                    cond_rows=cond_rows, cond_index=cidx)
        return flat_grad

    if ctx.n_importance > 0:
        if g_disp is not None:
            g_depth, g_acc = _disp_chain(g_disp, disp, acc, depth, g_depth, g_acc)
        g_depth0 = None
        if g_disp0 is not None:
            g_depth0, g_acc0 = _disp_chain(g_disp0, disp0, acc0, depth0, None, g_acc0)
        g_flat_f = grad_net(packed_f, z_f, raw_f, noise_f, g_rgb, g_depth, g_acc, g_rawf, ctx.n_f, tape_f)
        g_flat_c = grad_net(packed_c, z_c, raw_c, noise_c, g_rgb0, g_depth0, g_acc0, g_rawc, ctx.n_c, tape_c)
    else:
        # one network: rgb0 / disp0 / acc0 are copies of the maps
        def add(a, b):
            return b if a is None else (a if b is None else a + b)
        g_rgb, g_disp, g_acc = add(g_rgb, g_rgb0), add(g_disp, g_disp0), add(g_acc, g_acc0)
        if g_disp is not None:
            g_depth, g_acc = _disp_chain(g_disp, disp, acc, depth, g_depth, g_acc)
        g_flat_f = None
        g_flat_c = grad_net(packed_c, z_c, raw_c, noise_c, g_rgb, g_depth, g_acc, g_rawc, ctx.n_c, tape_c)
    return (g_flat_c, g_flat_f) + (None,) * (_N_INPUTS - 2)


render_rays_op.register_autograd(_backward, setup_context=_setup_context)

_PRECISION_NAMES = {v: k for k, v in ops.PRECISIONS.items()}
_OUT_NAMES = ["rgb", "disp", "acc", "depth", "rgb0", "disp0", "acc0", "z_std", "z_c", "z_f", "raw_c", "raw_f", "depth0"]
_validated_view_ids: Dict[tuple, bool] = {}


def _check_view_id(view_id: torch.Tensor, cidx: torch.Tensor, n_codes: int) -> None:
    """Range check of caller-supplied code indices (one device sync per distinct tensor version; the kernels also
    clamp, so a bad id can never read out of bounds -- but it is the caller's bug and is reported as one)."""
    key = (view_id.data_ptr(), view_id._version, view_id.numel(), str(view_id.device), n_codes)
    if _validated_view_ids.get(key):
        return
    lo, hi = (int(v) for v in torch.stack([cidx.min(), cidx.max()]).tolist())
    if lo < 0 or hi >= n_codes:
        raise ValueError(f"view_id must lie in [0, {n_codes}) (cond has {n_codes} rows); got range [{lo}, {hi}]")
    if len(_validated_view_ids) > 64:
        _validated_view_ids.clear()
    _validated_view_ids[key] = True



def _per_ray(v, R: int, device, name: str) -> torch.Tensor:
    if torch.is_tensor(v):
        t = v.to(device=device, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(R)
        if t.numel() != R:
            raise ValueError(f"{name} must be a float or have R={R} elements")
        return t.contiguous()
    return torch.full((R,), float(v), dtype=torch.float32, device=device)


def render_rays(model: NerfModel, rays_o: torch.Tensor, rays_d: torch.Tensor, near, far, N_samples: int,
                N_importance: int, cond: Optional[torch.Tensor] = None, *, view_id: Optional[torch.Tensor] = None,
                u_strat: Optional[torch.Tensor] = None, u_fine: Optional[torch.Tensor] = None, raw_noise=None,
                white_bkgd: bool = False, lindisp: bool = False, precision: str = "bf16",
                return_taps: bool = False, save_tape: Optional[bool] = None,
                fuse_composite: Optional[bool] = None) -> Dict[str, torch.Tensor]:
    """Volume-render a batch of rays (A.9).

    rays_o, rays_d: [R,3] CUDA fp32 (rays_d un-normalised).  near/far: floats or [R]/[R,1] tensors.
    cond: None | [256] | [R,256] | [V,256] with view_id[R] (A.8; requires a model built with cond=True).
    u_strat [R,N_samples] / u_fine [R,N_importance]: caller-supplied uniforms; None = deterministic
    (no jitter; u_fine = linspace(0,1,N_importance)).
    raw_noise: None, or the noise added to sigma_raw before the ReLU (A.5): a tensor [R,N_samples] when
    N_importance == 0, else a pair (coarse [R,N_samples], fine [R,N_samples+N_importance]); either entry may be None.
    save_tape: record the networks' activations during the forward so that backward() skips the recompute
    (bf16 path; ~5.4 KB per sample of HBM until backward).  None = automatically, when gradients are
    being recorded for the model's parameters.
    fuse_composite: run alpha compositing inside the bf16 network-query kernel wherever the sample
    count allows it; without return_taps raw[R,S,4] is then never written to HBM.  None = automatically, on the
    inference path (bf16, no tape, no gradient recorded for the parameters).
    Returns rgb[R,3], disp, acc, depth, rgb0, disp0, acc0, z_std (+ taps z_c, z_f, raw_c, raw_f, depth0).
    """
    if not rays_o.is_cuda:
        raise _lib.FnerfError("render_rays needs CUDA tensors (no CPU fallback)")
    dev = rays_o.device
    R = rays_o.shape[0]
    rays_o = rays_o.float().contiguous()
    rays_d = rays_d.float().contiguous()
    near_t, far_t = _per_ray(near, R, dev, "near"), _per_ray(far, R, dev, "far")
    t_vals = _linspace01(N_samples, dev)
    if u_strat is not None:
        u_strat = u_strat.float().contiguous()
        assert u_strat.shape == (R, N_samples)
    if N_importance > 0:
        if u_fine is None:
            u_fine = _linspace01(N_importance, dev)
        else:
            u_fine = u_fine.float().contiguous()
            assert u_fine.shape == (R, N_importance)
    noise_c = noise_f = None
    if raw_noise is not None:
        if torch.is_tensor(raw_noise):
            if N_importance > 0:
                raise ValueError("raw_noise must be a (coarse, fine) pair when N_importance > 0")
            raw_noise = (raw_noise, None)
        noise_c, noise_f = raw_noise
        if noise_c is not None:
            noise_c = noise_c.to(dev).float().contiguous()
            assert noise_c.shape == (R, N_samples), "coarse raw_noise must be [R, N_samples]"
        if noise_f is not None:
            noise_f = noise_f.to(dev).float().contiguous()
            assert N_importance > 0 and noise_f.shape == (R, N_samples + N_importance), "fine raw_noise must be [R, N_samples + N_importance]"
    cpc = cpf = cidx = None
    if cond is not None:
        if not model.cond:
            raise ValueError("cond given but the model was built without the conditioned layer 5")
        if cond.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("render_rays does not differentiate w.r.t. the garment codes (network inputs carry no "
                                      "gradient, SURVEY.md A.4); detach cond or train the codes outside this operator")
        cond = cond.to(dev).float().reshape(-1, 256)
        if view_id is not None:
            cidx = view_id.to(dev).to(torch.int32).contiguous()
            if cidx.numel() != R:
                raise ValueError(f"view_id must have R={R} elements")
            _check_view_id(view_id, cidx, cond.shape[0])
        elif cond.shape[0] not in (1, R):
            raise ValueError("cond must be [256], [R,256], or [V,256] with view_id")
        cpc = ops.cond_project(model.coarse.packed, cond)
        cpf = ops.cond_project(model.fine.packed, cond)
    elif model.cond:
        raise ValueError("model expects cond")
    can_tape = precision == "bf16"
    if save_tape is None:
        save_tape = can_tape and torch.is_grad_enabled() and (model.coarse.flat.requires_grad or model.fine.flat.requires_grad)
    elif save_tape and not can_tape:
        raise ValueError("save_tape needs precision='bf16'")
    needs_grad = torch.is_grad_enabled() and (model.coarse.flat.requires_grad or model.fine.flat.requires_grad)
    if fuse_composite is None:
        fuse_composite = precision == "bf16" and not save_tape and not needs_grad
    elif fuse_composite and (precision != "bf16" or save_tape):
        raise ValueError("fuse_composite needs precision='bf16' and no tape")
    outs = render_rays_op(model.coarse.flat, model.fine.flat, model.coarse.packed, model.fine.packed, rays_o, rays_d,
                          near_t, far_t, t_vals, u_strat, u_fine if N_importance > 0 else None, noise_c, noise_f, cpc, cpf, cidx,
                          cond if cond is not None else None, int(N_importance), bool(white_bkgd), bool(lindisp),
                          ops.PRECISIONS[precision], bool(save_tape), bool(fuse_composite),
                          bool(return_taps or needs_grad))
    n = len(_OUT_NAMES) if return_taps else 8
    return {k: v for k, v in zip(_OUT_NAMES[:n], outs[:n])}


def render_image(model: NerfModel, rays_o: torch.Tensor, rays_d: torch.Tensor, near, far, N_samples: int,
                 N_importance: int, cond=None, *, chunk: int = 1 << 16, **kw) -> Dict[str, torch.Tensor]:
    """Full-frame render: chunk the flattened ray list through render_rays."""
    keys = ("rgb", "disp", "acc", "depth", "rgb0", "disp0", "acc0", "z_std")
    parts = {k: [] for k in keys}
    R = rays_o.shape[0]
    per_ray = {k: kw.pop(k) for k in ("u_strat", "u_fine", "view_id") if k in kw}
    noise = kw.pop("raw_noise", None)
    if torch.is_tensor(noise):
        noise = (noise, None)
    with torch.no_grad():
        for s in range(0, R, chunk):
            sl = slice(s, min(s + chunk, R))
            extra = {k: (v[sl] if v is not None else None) for k, v in per_ray.items()}
            c = cond[sl] if (cond is not None and cond.dim() == 2 and cond.shape[0] == R) else cond
            if noise is not None:
                extra["raw_noise"] = tuple(None if t is None else t[sl] for t in noise)
                if N_importance == 0:
                    extra["raw_noise"] = extra["raw_noise"][0]
            nr, fr = (v[sl] if (torch.is_tensor(v) and v.numel() == R) else v for v in (near, far))
            out = render_rays(model, rays_o[sl], rays_d[sl], nr, fr, N_samples, N_importance, c, **extra, **kw)
            for k in keys:
                parts[k].append(out[k])
    return {k: torch.cat(v, 0) for k, v in parts.items()}
