"""CPU oracle for the NeRF render/train hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``fashion_nerf_b200``) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference mount holds no code, tests or golden vectors
(/root/reference/README.md:1-2 is the whole tree), so nothing of the reference's can
pin this oracle.  It is a plain fp32 PyTorch restatement of the canonical NeRF
equations as written down in SURVEY.md Appendix A (sections cited per function),
validated by the closed-form self-tests in tests/test_oracle.py.

Reproducibility rules, so that the CUDA kernels can be
bit-exact on sample positions and bin indices:
  * only separately rounded elementwise ops (no lerp / addcmul / fused multiply-add);
  * the pdf normaliser and the CDF are accumulated in fp64 and rounded once per
    output (exact, hence order independent, because of the +1e-5 floor);
  * ``t_vals`` / deterministic ``u`` are ``torch.linspace`` tensors handed to both
    sides (``linspace(0,1,n)[i] != i/(n-1)`` in fp32).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

L_XYZ = 10   # A.3: 63-d encoding of positions
L_DIR = 4    # A.3: 27-d encoding of view directions
PE_XYZ = 3 + 6 * L_XYZ
PE_DIR = 3 + 6 * L_DIR
WIDTH = 256
DEPTH = 8
SKIP = 4
COND_DIM = 256


# --------------------------------------------------------------------------- A.1
def _norm3(d: torch.Tensor) -> torch.Tensor:
    """|d| with a pinned evaluation order, (x*x + y*y) + z*z in fp32, and a CORRECTLY ROUNDED square
    root (evaluated in fp64, rounded once).  torch.sqrt on an AVX-512 CPU is 1 ulp off the IEEE
    result in ~0.2 % of inputs (measured on the GPU box), so it cannot define a bit-exact contract;
    CUDA's sqrt.rn.f32 is the IEEE result."""
    n2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return torch.sqrt(n2.double()).float()


def ray_setup(rays_d: torch.Tensor):
    """viewdirs = d/|d|, dnorm = |d| (rays_d itself stays un-normalised)."""
    dnorm = _norm3(rays_d)
    return rays_d / dnorm[:, None], dnorm


ray_setup_exact = ray_setup


# --------------------------------------------------------------------------- A.2
def stratified(near: torch.Tensor, far: torch.Tensor, t_vals: torch.Tensor,
               u_strat: Optional[torch.Tensor] = None, lindisp: bool = False) -> torch.Tensor:
    """SURVEY.md A.2.  near/far: [R]; t_vals: [N]; u_strat: [R,N] or None.  Returns z [R,N]."""
    near = near[:, None]
    far = far[:, None]
    t = t_vals[None, :]
    if not lindisp:
        z = near * (1.0 - t) + far * t
    else:
        z = 1.0 / ((1.0 / near) * (1.0 - t) + (1.0 / far) * t)
    if u_strat is not None:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * u_strat
    return z


# --------------------------------------------------------------------------- A.3
def posenc(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], no pi."""
    out = [x]
    for k in range(L):
        f = float(2 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


# --------------------------------------------------------------------------- A.4 / A.8
LAYER_NAMES = [f"pts_linears.{i}" for i in range(DEPTH)] + [
    "alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"]


def layer_shapes(cond: bool = False) -> Dict[str, tuple]:
    """(out, in) of every nn.Linear of one network (A.4; A.8 widens layer 5 by COND_DIM)."""
    s = {}
    for i in range(DEPTH):
        if i == 0:
            k = PE_XYZ
        elif i == SKIP + 1:
            k = PE_XYZ + (COND_DIM if cond else 0) + WIDTH
        else:
            k = WIDTH
        s[f"pts_linears.{i}"] = (WIDTH, k)
    s["alpha_linear"] = (1, WIDTH)
    s["feature_linear"] = (WIDTH, WIDTH)
    s["views_linears.0"] = (WIDTH // 2, WIDTH + PE_DIR)
    s["rgb_linear"] = (3, WIDTH // 2)
    return s


def init_params(seed: int, cond: bool = False) -> Dict[str, torch.Tensor]:
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for W and b),
    drawn from a private generator so it does not depend on global RNG state."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, (o, i) in layer_shapes(cond).items():
        bound = 1.0 / math.sqrt(i)
        p[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    return p


def mlp_forward(p: Dict[str, torch.Tensor], pe: torch.Tensor, pe_dir: torch.Tensor,
                cond: Optional[torch.Tensor] = None, *, bf16: bool = False) -> torch.Tensor:
    """SURVEY.md A.4 (+A.8).  pe [M,63], pe_dir [M,27], cond [M,256] or None -> raw [M,4] (rgb, sigma).

    ``bf16=True`` emulates the tensor-core path's rounding points (inputs, weights and
    inter-layer activations rounded to bf16, fp32 accumulate, fp32 biases) -- used only to
    explain error budgets in tests, never as the parity oracle.
    """
    def rnd(t):
        return t.to(torch.bfloat16).to(torch.float32) if bf16 else t

    def lin(name, x):
        return x @ rnd(p[name + ".weight"]).t() + p[name + ".bias"]

    pe = rnd(pe)
    pe_dir = rnd(pe_dir)
    h = pe
    for i in range(DEPTH):
        h = rnd(torch.relu(lin(f"pts_linears.{i}", h)))
        if i == SKIP:
            if cond is not None:
                h = torch.cat([pe, rnd(cond), h], -1)
            else:
                h = torch.cat([pe, h], -1)
    sigma = lin("alpha_linear", h)
    feat = rnd(lin("feature_linear", h))
    hv = rnd(torch.relu(lin("views_linears.0", torch.cat([feat, pe_dir], -1))))
    rgb = lin("rgb_linear", hv)
    return torch.cat([rgb, sigma], -1)


def run_network(p, pts: torch.Tensor, viewdirs: torch.Tensor, cond_rows: Optional[torch.Tensor] = None,
                chunk: int = 65536, bf16: bool = False) -> torch.Tensor:
    """pts [R,S,3], viewdirs [R,3] (+ cond_rows [R,256]) -> raw [R,S,4]; chunked to bound memory (3.4)."""
    R, S, _ = pts.shape
    flat = pts.reshape(-1, 3)
    dirs = viewdirs[:, None, :].expand(R, S, 3).reshape(-1, 3)
    cflat = None if cond_rows is None else cond_rows[:, None, :].expand(R, S, COND_DIM).reshape(-1, COND_DIM)
    outs = []
    for s in range(0, flat.shape[0], chunk):
        pe = posenc(flat[s:s + chunk], L_XYZ)
        ped = posenc(dirs[s:s + chunk], L_DIR)
        c = None if cflat is None else cflat[s:s + chunk]
        outs.append(mlp_forward(p, pe, ped, c, bf16=bf16))
    return torch.cat(outs, 0).reshape(R, S, 4)


# --------------------------------------------------------------------------- A.5
def raw2outputs(raw: torch.Tensor, z: torch.Tensor, dnorm: torch.Tensor,
                white_bkgd: bool = False, raw_noise: Optional[torch.Tensor] = None):
    """SURVEY.md A.5.  raw [R,S,4], z [R,S], dnorm [R] -> dict(rgb, disp, acc, depth, weights)."""
    dists = z[:, 1:] - z[:, :-1]
    dists = torch.cat([dists, torch.full_like(z[:, :1], 1e10)], -1)
    dists = dists * dnorm[:, None]
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3]
    if raw_noise is not None:
        sigma = sigma + raw_noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * dists)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * T
    rgb_map = (weights[..., None] * rgb).sum(-2)
    depth_map = (weights * z).sum(-1)
    acc_map = weights.sum(-1)
    disp_map = 1.0 / torch.maximum(torch.full_like(depth_map, 1e-10), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])
    return {"rgb": rgb_map, "disp": disp_map, "acc": acc_map, "depth": depth_map, "weights": weights}


# --------------------------------------------------------------------------- A.6
def composite_bwd(raw: torch.Tensor, z: torch.Tensor, dnorm: torch.Tensor,
                  g_rgb: torch.Tensor, g_depth: Optional[torch.Tensor] = None,
                  g_acc: Optional[torch.Tensor] = None, white_bkgd: bool = False,
                  raw_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """closed-form dL/draw [R,S,4] given dL/d(rgb_map, depth_map, acc_map).

    (disp_map carries no gradient in the training loss, A.10.)
    """
    R, S, _ = raw.shape
    gd = torch.zeros(R, dtype=raw.dtype) if g_depth is None else g_depth
    ga = torch.zeros(R, dtype=raw.dtype) if g_acc is None else g_acc
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1) * dnorm[:, None]
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3]
    if raw_noise is not None:
        sigma = sigma + raw_noise
    alpha = 1.0 - torch.exp(-torch.relu(sigma) * dists)
    one_m = 1.0 - alpha + 1e-10
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), one_m], -1), -1)[:, :-1]
    w = alpha * T
    v = (g_rgb[:, None, :] * rgb).sum(-1) + gd[:, None] * z + ga[:, None]
    if white_bkgd:
        v = v - g_rgb.sum(-1)[:, None]
    wv = w * v
    suffix = torch.flip(torch.cumsum(torch.flip(wv, [-1]), -1), [-1]) - wv     # sum_{k>i}
    g_alpha = T * v - suffix / one_m
    g_sigma = (sigma > 0).to(raw.dtype) * dists * (1.0 - alpha) * g_alpha
    g_rgb_raw = w[..., None] * g_rgb[:, None, :] * rgb * (1.0 - rgb)
    return torch.cat([g_rgb_raw, g_sigma[..., None]], -1)


# --------------------------------------------------------------------------- A.7
def sample_pdf(z_c: torch.Tensor, weights_c: torch.Tensor, u: torch.Tensor):
    """SURVEY.md A.7.  z_c [R,Nc], weights_c [R,Nc], u [R,Nf] -> dict(inds, below, above, z_samples, z_f, z_std)."""
    Nc = z_c.shape[-1]
    bins = 0.5 * (z_c[:, 1:] + z_c[:, :-1])                      # [R, Nc-1]
    w = weights_c[:, 1:-1] + 1e-5                                 # [R, Nc-2]
    norm = w.double().sum(-1, keepdim=True).float()               # exactly rounded normaliser (H1)
    pdf = w / norm
    cdf = torch.cumsum(pdf.double(), -1).float()                  # fp64 accumulate, round per output (H1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)      # [R, Nc-1]
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                 # count of cdf <= u
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=Nc - 2)
    cdf_b = torch.gather(cdf, 1, below)
    cdf_a = torch.gather(cdf, 1, above)
    bin_b = torch.gather(bins, 1, below)
    bin_a = torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    z_samples = bin_b + t * (bin_a - bin_b)
    z_f, _ = torch.sort(torch.cat([z_c, z_samples], -1), -1)
    z_std = torch.std(z_samples, -1, unbiased=False)
    return {"inds": inds, "below": below, "above": above, "z_samples": z_samples, "z_f": z_f,
            "z_std": z_std, "cdf": cdf, "bins": bins}


# --------------------------------------------------------------------------- A.9
def render_rays(params_c, params_f, rays_o, rays_d, near, far, N_samples, N_importance,
                cond_rows=None, *, u_strat=None, u_fine=None, raw_noise=None, white_bkgd=False, lindisp=False,
                bf16=False, z_f_override=None, return_extras=False):
    """coarse -> composite -> importance -> fine -> composite.

    near/far: floats or [R] tensors.  cond_rows: [R,256] or None (already gathered per ray).
    ``z_f_override`` teacher-forces the fine depths (H6).  ``raw_noise``: None or a pair
    (coarse [R,N_samples], fine [R,N_samples+N_importance]) added to sigma_raw before the ReLU (A.5).  All fp32, CPU.
    """
    noise_c, noise_f = (None, None) if raw_noise is None else raw_noise
    R = rays_o.shape[0]
    near = torch.full((R,), float(near)) if not torch.is_tensor(near) else near.reshape(R).float()
    far = torch.full((R,), float(far)) if not torch.is_tensor(far) else far.reshape(R).float()
    viewdirs, dnorm = ray_setup_exact(rays_d)
    t_vals = torch.linspace(0.0, 1.0, N_samples)
    z_c = stratified(near, far, t_vals, u_strat, lindisp)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_c[:, :, None]
    raw_c = run_network(params_c, pts, viewdirs, cond_rows, bf16=bf16)
    out_c = raw2outputs(raw_c, z_c, dnorm, white_bkgd, noise_c)
    res = {"rgb0": out_c["rgb"], "disp0": out_c["disp"], "acc0": out_c["acc"], "depth0": out_c["depth"]}
    extras = {"z_c": z_c, "raw_c": raw_c, "weights_c": out_c["weights"], "dnorm": dnorm, "viewdirs": viewdirs}
    if N_importance > 0:
        if u_fine is None:
            u_fine = torch.linspace(0.0, 1.0, N_importance)[None, :].expand(R, N_importance)
        sp = sample_pdf(z_c, out_c["weights"], u_fine)
        z_f = sp["z_f"] if z_f_override is None else z_f_override
        pts = rays_o[:, None, :] + rays_d[:, None, :] * z_f[:, :, None]
        raw_f = run_network(params_f, pts, viewdirs, cond_rows, bf16=bf16)
        out_f = raw2outputs(raw_f, z_f, dnorm, white_bkgd, noise_f)
        res.update({"rgb": out_f["rgb"], "disp": out_f["disp"], "acc": out_f["acc"],
                    "depth": out_f["depth"], "z_std": sp["z_std"]})
        extras.update({"z_f": z_f, "raw_f": raw_f, "weights_f": out_f["weights"], "inds": sp["inds"],
                       "z_samples": sp["z_samples"]})
    else:
        res.update({"rgb": out_c["rgb"], "disp": out_c["disp"], "acc": out_c["acc"],
                    "depth": out_c["depth"], "z_std": torch.zeros(R)})
    if return_extras:
        res["extras"] = extras
    return res


# --------------------------------------------------------------------------- A.10
def loss_and_grads(params_c, params_f, rays_o, rays_d, near, far, N_samples, N_importance, target,
                   cond_rows=None, *, u_strat=None, u_fine=None, white_bkgd=False):
    """loss = mse(rgb, tgt) + mse(rgb0, tgt); grads via torch.autograd through the
    fp32 oracle (z_samples detached, A.7).  Returns (loss, grads_c, grads_f) with fresh leaf copies."""
    pc = {k: v.clone().requires_grad_(True) for k, v in params_c.items()}
    pf = {k: v.clone().requires_grad_(True) for k, v in params_f.items()}
    R = rays_o.shape[0]
    near_t = torch.full((R,), float(near)) if not torch.is_tensor(near) else near.reshape(R).float()
    far_t = torch.full((R,), float(far)) if not torch.is_tensor(far) else far.reshape(R).float()
    viewdirs, dnorm = ray_setup_exact(rays_d)
    t_vals = torch.linspace(0.0, 1.0, N_samples)
    z_c = stratified(near_t, far_t, t_vals, u_strat)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_c[:, :, None]
    raw_c = run_network(pc, pts, viewdirs, cond_rows)
    out_c = raw2outputs(raw_c, z_c, dnorm, white_bkgd)
    if u_fine is None:
        u_fine = torch.linspace(0.0, 1.0, N_importance)[None, :].expand(R, N_importance)
    with torch.no_grad():
        z_f = sample_pdf(z_c, out_c["weights"].detach(), u_fine)["z_f"]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_f[:, :, None]
    raw_f = run_network(pf, pts, viewdirs, cond_rows)
    out_f = raw2outputs(raw_f, z_f, dnorm, white_bkgd)
    loss = ((out_f["rgb"] - target) ** 2).mean() + ((out_c["rgb"] - target) ** 2).mean()
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in pc.items()}, {k: v.grad for k, v in pf.items()}


def adam_step(params, grads, state, lr=5e-4, b1=0.9, b2=0.999, eps=1e-8):
    """Plain Adam (torch defaults; A.10 pins eps=1e-8).  state: dict(step, m{}, v{})."""
    state["step"] = state.get("step", 0) + 1
    t = state["step"]
    for k in params:
        m = state.setdefault("m", {}).setdefault(k, torch.zeros_like(params[k]))
        v = state.setdefault("v", {}).setdefault(k, torch.zeros_like(params[k]))
        m.mul_(b1).add_(grads[k], alpha=1 - b1)
        v.mul_(b2).addcmul_(grads[k], grads[k], value=1 - b2)
        mhat = m / (1 - b1 ** t)
        vhat = v / (1 - b2 ** t)
        params[k] = params[k] - lr * mhat / (vhat.sqrt() + eps)
    return params


# --------------------------------------------------------------------------- synthetic inputs (8d)
def pinhole_rays(H: int, W: int, view: int = 0, n_views: int = 1):
    """SURVEY.md 8(d) camera: origin (0,0,4) looking down -z, horizontal FOV 0.6911 rad, un-normalised
    directions; view v rotates by 2*pi*v/V about +y.  Returns rays_o, rays_d  [H*W,3] fp32."""
    f = 0.5 * W / math.tan(0.34555)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32),
                          indexing="ij")
    d = torch.stack([(i - W * 0.5) / f, -(j - H * 0.5) / f, -torch.ones_like(i)], -1).reshape(-1, 3)
    o = torch.tensor([0.0, 0.0, 4.0]).expand_as(d).clone()
    if n_views > 1 and view != 0:
        a = 2.0 * math.pi * view / n_views
        c, s = math.cos(a), math.sin(a)
        Rm = torch.tensor([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=torch.float32)
        d = d @ Rm.t()
        o = o @ Rm.t()
    return o.contiguous(), d.contiguous()


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = ((a - b) ** 2).mean().item()
    return -10.0 * math.log10(max(mse, 1e-20))
