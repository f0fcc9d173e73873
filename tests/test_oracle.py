"""Closed-form self-tests of the CPU oracle ("Oracle self-tests").

The reference ships no tests or golden vectors (parity unpinned), so the oracle is pinned against
analytic answers and against torch.autograd / fp64 re-evaluations of the same equations."""
import math

import pytest
import torch

from oracle import nerf_oracle as O


def test_param_count_matches_survey():
    p = O.init_params(0)
    assert sum(v.numel() for v in p.values()) == 595_844
    pc = O.init_params(0, cond=True)
    assert pc["pts_linears.5.weight"].shape == (256, 63 + 256 + 256)  # A.8


def test_stratified_no_jitter_is_linspace():
    near, far = torch.tensor([2.0, 1.0]), torch.tensor([6.0, 3.0])
    t = torch.linspace(0, 1, 64)
    z = O.stratified(near, far, t)
    assert z.shape == (2, 64)
    assert torch.equal(z[:, 0], near) and torch.equal(z[:, -1], far)
    assert (z[:, 1:] > z[:, :-1]).all()


def test_stratified_jitter_stays_in_bins():
    torch.manual_seed(0)
    near, far = torch.full((16,), 2.0), torch.full((16,), 6.0)
    t = torch.linspace(0, 1, 64)
    z0 = O.stratified(near, far, t)
    mids = 0.5 * (z0[:, 1:] + z0[:, :-1])
    lo = torch.cat([z0[:, :1], mids], -1)
    hi = torch.cat([mids, z0[:, -1:]], -1)
    z = O.stratified(near, far, t, torch.rand(16, 64))
    assert (z >= lo).all() and (z <= hi).all()
    assert torch.equal(O.stratified(near, far, t, torch.zeros(16, 64)), lo)


def test_stratified_lindisp():
    near, far = torch.tensor([2.0]), torch.tensor([6.0])
    z = O.stratified(near, far, torch.linspace(0, 1, 5), lindisp=True)
    assert torch.allclose(1.0 / z, torch.linspace(0.5, 1 / 6, 5)[None], atol=1e-6)


def test_posenc_known_answers():
    x = torch.zeros(1, 3)
    pe = O.posenc(x, 10)
    assert pe.shape == (1, 63)
    expect = torch.cat([torch.zeros(3)] + [torch.cat([torch.zeros(3), torch.ones(3)]) for _ in range(10)])
    assert torch.equal(pe[0], expect)
    # x = pi/2^k: sin(2^k x) = 0, cos(2^k x) = -1; octave k-1: sin = 1, cos = 0
    k = 3
    x = torch.full((1, 3), math.pi / 2 ** k)
    pe = O.posenc(x, 4)
    assert torch.allclose(pe[0, 3 + 6 * k:3 + 6 * k + 3], torch.zeros(3), atol=1e-6)
    assert torch.allclose(pe[0, 3 + 6 * k + 3:3 + 6 * k + 6], -torch.ones(3), atol=1e-6)
    assert torch.allclose(pe[0, 3 + 6 * (k - 1):3 + 6 * (k - 1) + 3], torch.ones(3), atol=1e-6)
    assert O.posenc(torch.rand(5, 3), 4).shape == (5, 27)


def test_mlp_matches_nn_sequential_fp64():
    """A.4 against an independent nn.Module re-statement evaluated in fp64."""
    import torch.nn as nn
    p = O.init_params(3)
    torch.manual_seed(0)
    pe, ped = torch.randn(32, 63), torch.randn(32, 27)

    class Ref(nn.Module):
        def __init__(self):
            super().__init__()
            self.pts = nn.ModuleList([nn.Linear(63, 256)] + [nn.Linear(256 if i != 4 else 319, 256) for i in range(7)])
            self.alpha, self.feat = nn.Linear(256, 1), nn.Linear(256, 256)
            self.views, self.rgb = nn.Linear(283, 128), nn.Linear(128, 3)

        def forward(self, x, d):
            h = x
            for i, l in enumerate(self.pts):
                h = torch.relu(l(h))
                if i == 4:
                    h = torch.cat([x, h], -1)
            a = self.alpha(h)
            hv = torch.relu(self.views(torch.cat([self.feat(h), d], -1)))
            return torch.cat([self.rgb(hv), a], -1)

    m = Ref().double()
    with torch.no_grad():
        for i, l in enumerate(m.pts):
            l.weight.copy_(p[f"pts_linears.{i}.weight"]); l.bias.copy_(p[f"pts_linears.{i}.bias"])
        for mod, name in [(m.alpha, "alpha_linear"), (m.feat, "feature_linear"), (m.views, "views_linears.0"),
                          (m.rgb, "rgb_linear")]:
            mod.weight.copy_(p[name + ".weight"]); mod.bias.copy_(p[name + ".bias"])
        ref = m(pe.double(), ped.double())
    got = O.mlp_forward(p, pe, ped)
    assert got.shape == (32, 4)
    assert (got.double() - ref).abs().max() < 1e-5


def test_mlp_cond_equals_dense_layer5():
    p = O.init_params(5, cond=True)
    torch.manual_seed(1)
    pe, ped, c = torch.randn(8, 63), torch.randn(8, 27), torch.randn(8, 256)
    raw = O.mlp_forward(p, pe, ped, c)
    # zero code == dropping the cond columns
    p0 = {k: v.clone() for k, v in p.items()}
    raw0 = O.mlp_forward(p0, pe, ped, torch.zeros(8, 256))
    p_nc = O.init_params(5, cond=False)
    for k in p_nc:
        p_nc[k] = p[k].clone()
    w5 = p["pts_linears.5.weight"]
    p_nc["pts_linears.5.weight"] = torch.cat([w5[:, :63], w5[:, 319:]], -1)
    assert torch.allclose(raw0, O.mlp_forward(p_nc, pe, ped), atol=1e-6)
    assert not torch.allclose(raw, raw0)


def test_compositing_constant_sigma_closed_form():
    """T_i = exp(-sigma * sum_{k<i} delta_k) for constant sigma; weights telescope to 1 - T_end."""
    R, S = 3, 48
    z = torch.linspace(2, 6, S)[None].repeat(R, 1)
    raw = torch.zeros(R, S, 4)
    sigma = 0.7
    raw[..., 3] = sigma
    raw[..., :3] = torch.tensor([0.3, -0.2, 1.1])
    dn = torch.tensor([1.0, 1.5, 2.0])
    out = O.raw2outputs(raw, z, dn)
    delta = (z[:, 1:] - z[:, :-1]) * dn[:, None]
    T = torch.exp(-sigma * torch.cat([torch.zeros(R, 1), torch.cumsum(delta, -1)], -1))
    alpha = 1 - torch.exp(-sigma * torch.cat([delta, torch.full((R, 1), 1e10)], -1))
    assert torch.allclose(out["weights"], alpha * T, atol=2e-6)
    assert torch.allclose(out["acc"], torch.ones(R), atol=1e-5)       # last sample is opaque
    assert torch.allclose(out["rgb"], torch.sigmoid(torch.tensor([0.3, -0.2, 1.1]))[None].expand(R, 3), atol=1e-5)


def test_compositing_empty_space_and_white_bkgd():
    R, S = 2, 16
    z = torch.linspace(2, 6, S)[None].repeat(R, 1)
    raw = torch.zeros(R, S, 4)
    raw[..., 3] = -1.0                                                # sigma <= 0 everywhere
    out = O.raw2outputs(raw, z, torch.ones(R), white_bkgd=True)
    assert torch.equal(out["weights"], torch.zeros(R, S))
    assert torch.equal(out["rgb"], torch.ones(R, 3))
    assert torch.isnan(out["disp"]).all()                             # 0/0 propagates like torch.max


@pytest.mark.parametrize("white", [False, True])
def test_composite_bwd_matches_autograd_fp64(white):
    """A.6 closed form vs torch.autograd of A.5, fp64 (SURVEY.md probe: 5.6e-16)."""
    torch.manual_seed(0)
    R, S = 64, 48
    raw = torch.randn(R, S, 4, dtype=torch.float64, requires_grad=True)
    z = torch.sort(torch.rand(R, S, dtype=torch.float64) * 4 + 2, -1)[0]
    dn = torch.rand(R, dtype=torch.float64) + 1
    g_rgb, g_d, g_a = (torch.randn(R, 3, dtype=torch.float64), torch.randn(R, dtype=torch.float64),
                       torch.randn(R, dtype=torch.float64))
    out = O.raw2outputs(raw, z, dn, white)
    loss = (out["rgb"] * g_rgb).sum() + (out["depth"] * g_d).sum() + (out["acc"] * g_a).sum()
    (ref,) = torch.autograd.grad(loss, raw)
    got = O.composite_bwd(raw.detach(), z, dn, g_rgb, g_d, g_a, white)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max() < 1e-12


def test_composite_bwd_fp32_last_sample_finite():
    torch.manual_seed(1)
    raw = torch.randn(8, 32, 4) * 3
    z = torch.sort(torch.rand(8, 32) * 4 + 2, -1)[0]
    g = O.composite_bwd(raw, z, torch.ones(8), torch.randn(8, 3))
    assert torch.isfinite(g).all()


def test_sample_pdf_searchsorted_semantics():
    cdf = torch.tensor([[0.0, 0.25, 0.25, 1.0]])
    u = torch.tensor([[0.0, 0.25, 0.3, 1.0]])
    assert torch.searchsorted(cdf, u, right=True).tolist() == [[1, 3, 3, 4]]      # SURVEY.md A.7 probe


def test_sample_pdf_uniform_weights_is_piecewise_linear():
    """Uniform weights => cdf linear in bin index => z_samples = bins[0] + u*(bins[-1]-bins[0])."""
    Nc, Nf = 64, 128
    z_c = torch.linspace(2, 6, Nc)[None]
    w = torch.full((1, Nc), 0.01)
    u = torch.linspace(0, 1, Nf)[None]
    sp = O.sample_pdf(z_c, w, u)
    bins = 0.5 * (z_c[:, 1:] + z_c[:, :-1])
    expect = bins[:, :1] + u * (bins[:, -1:] - bins[:, :1])
    assert torch.allclose(sp["z_samples"], expect, atol=2e-5)
    assert sp["z_f"].shape == (1, Nc + Nf)
    assert (sp["z_f"][:, 1:] >= sp["z_f"][:, :-1]).all()
    assert sp["inds"].min() >= 1 and sp["inds"].max() <= Nc - 1


def test_sample_pdf_edge_u_and_zero_weights():
    Nc = 16
    z_c = torch.linspace(2, 6, Nc)[None]
    w = torch.zeros(1, Nc)                                            # pdf uniform via the +1e-5 floor
    u = torch.tensor([[0.0, 0.5, 0.99999994, 1.0, 1.5]])
    sp = O.sample_pdf(z_c, w, u)
    assert torch.isfinite(sp["z_samples"]).all()
    assert sp["inds"][0, 0] == 1 and sp["inds"][0, -1] == Nc - 1      # u beyond cdf[-1] lands on index n (clamped)
    bins = sp["bins"]
    assert sp["z_samples"].min() >= bins.min() - 1e-6


def test_sample_pdf_concentrates_on_heavy_bin():
    Nc, Nf = 64, 128
    z_c = torch.linspace(2, 6, Nc)[None]
    w = torch.zeros(1, Nc)
    w[0, 30] = 1.0
    torch.manual_seed(0)
    sp = O.sample_pdf(z_c, w, torch.rand(1, Nf))
    bins = sp["bins"]
    inside = (sp["z_samples"] >= bins[0, 29]) & (sp["z_samples"] <= bins[0, 30])
    assert inside.float().mean() > 0.95


def test_render_rays_shapes_and_determinism():
    pc, pf = O.init_params(0), O.init_params(1)
    o, d = O.pinhole_rays(4, 4)
    a = O.render_rays(pc, pf, o, d, 2.0, 6.0, 16, 8)
    b = O.render_rays(pc, pf, o, d, 2.0, 6.0, 16, 8)
    for k in ("rgb", "disp", "acc", "depth", "rgb0", "disp0", "acc0", "z_std"):
        assert a[k].shape[0] == 16 and torch.allclose(a[k], b[k], rtol=0, atol=0, equal_nan=True)
    assert a["rgb"].shape == (16, 3)
    assert (d.norm(dim=-1) >= 1.0).all() and (d.norm(dim=-1) > 1.0).any()                               # un-normalised directions (8d)


def test_loss_grads_finite_and_adam_moves():
    pc, pf = O.init_params(0), O.init_params(1)
    o, d = O.pinhole_rays(2, 4)
    tgt = torch.rand(8, 3, generator=torch.Generator().manual_seed(100))
    loss, gc, gf = O.loss_and_grads(pc, pf, o, d, 2.0, 6.0, 8, 8, tgt)
    assert torch.isfinite(loss)
    assert all(torch.isfinite(g).all() for g in gc.values()) and all(torch.isfinite(g).all() for g in gf.values())
    st = {}
    new = O.adam_step(dict(pc), gc, st)
    assert not torch.equal(new["rgb_linear.bias"], pc["rgb_linear.bias"])


def test_bf16_emulation_is_close_to_fp32():
    p = O.init_params(0)
    torch.manual_seed(0)
    x = torch.rand(64, 3) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(64, 3), dim=-1)
    pe, ped = O.posenc(x, 10), O.posenc(dirs, 4)
    a, b = O.mlp_forward(p, pe, ped), O.mlp_forward(p, pe, ped, bf16=True)
    assert (a - b).abs().max() < 2e-2
