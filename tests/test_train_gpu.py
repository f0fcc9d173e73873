"""Backward / training-step parity (A.6 + A.4 backward + A.10) against torch.autograd through the fp32
CPU oracle on identical inputs.  Gradients flow to both networks' parameters; sample positions are
detached (A.7)."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F(cuda_device):
    import fashion_nerf_b200 as f
    f.load_library()
    return f


def _flat_grads(F, g, cond=False):
    return F.flatten_state_dict({k: v for k, v in g.items()}, cond)


def _rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("cond", [False, True])
def test_mlp_bwd_matches_autograd(F, cuda_device, cond):
    """fnerf_mlp_bwd vs autograd of the oracle MLP for a random upstream gradient on raw."""
    dev = cuda_device
    R, S, V = 40, 24, 3
    g = torch.Generator().manual_seed(21)
    o = torch.rand(R, 3, generator=g) * 2 - 1
    d = torch.randn(R, 3, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    vd, _ = O.ray_setup(d)
    p = O.init_params(4, cond=cond)
    codes = 0.25 * torch.randn(V, 256, generator=g) if cond else None
    vid = torch.randint(0, V, (R,), generator=g) if cond else None
    g_raw = torch.randn(R, S, 4, generator=g)
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    raw = O.run_network(pr, pts, vd, codes[vid] if cond else None)
    (raw * g_raw).sum().backward()
    ref = _flat_grads(F, {k: v.grad for k, v in pr.items()}, cond)

    net = F.NerfNetwork.from_state_dict(p, dev, cond=cond)
    flat_grad = torch.zeros(net.flat.numel(), device=dev)
    F.ops.mlp_bwd(net.packed, o.to(dev), d.to(dev), vd.to(dev), z.to(dev), g_raw.to(dev), flat_grad,
                  cond_rows=codes.to(dev) if cond else None, cond_index=vid.to(dev) if cond else None)
    got = flat_grad.cpu()
    assert torch.isfinite(got).all()
    sd_ref, sd_got = F.unflatten(ref, cond), F.unflatten(got, cond)
    for k in sd_ref:
        err = _rel_err(sd_got[k], sd_ref[k])
        assert err <= 2e-4, (k, err)
    # accumulation semantics: a second call adds
    F.ops.mlp_bwd(net.packed, o.to(dev), d.to(dev), vd.to(dev), z.to(dev), g_raw.to(dev), flat_grad,
                  cond_rows=codes.to(dev) if cond else None, cond_index=vid.to(dev) if cond else None)
    assert _rel_err(flat_grad.cpu(), 2 * ref) <= 2e-4


def test_mlp_bwd_multi_chunk(F, cuda_device):
    """More samples than one workspace chunk (32768): linearity check against two half calls."""
    dev = cuda_device
    R, S = 700, 64                                                   # 44800 samples -> 2 chunks
    g = torch.Generator().manual_seed(22)
    o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev)
    d = torch.randn(R, 3, generator=g).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
    g_raw = torch.randn(R, S, 4, generator=g).to(dev)
    net = F.NerfNetwork.random(4, dev)
    vd, _ = F.ops.ray_setup(d)
    full = torch.zeros(net.flat.numel(), device=dev)
    F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, full)
    halves = torch.zeros_like(full)
    for sl in (slice(0, 350), slice(350, 700)):
        F.ops.mlp_bwd(net.packed, o[sl], d[sl], vd[sl], z[sl], g_raw[sl], halves)
    assert _rel_err(full, halves) <= 1e-4


def test_render_rays_autograd_matches_oracle(F, cuda_device):
    """Full A.10 gradient: loss = mse(rgb) + mse(rgb0) through render_rays (fp32 forward)."""
    dev = cuda_device
    pc, pf = O.init_params(0), O.init_params(1)
    for p in (pc, pf):
        p["alpha_linear.bias"] += 0.1                                 # keep density away from the relu/step kink
    o, d = O.pinhole_rays(12, 16)
    R, Nc, Nf = o.shape[0], 32, 32
    g = torch.Generator().manual_seed(0)
    u_s, u_f = torch.rand(R, Nc, generator=g), torch.rand(R, Nf, generator=g)
    tgt = torch.rand(R, 3, generator=torch.Generator().manual_seed(100))
    loss_ref, gc, gf = O.loss_and_grads(pc, pf, o, d, 2.0, 6.0, Nc, Nf, tgt, u_strat=u_s, u_fine=u_f)
    model = F.NerfModel(F.NerfNetwork.from_state_dict(pc, dev), F.NerfNetwork.from_state_dict(pf, dev))
    model.coarse.flat.requires_grad_(True)
    model.fine.flat.requires_grad_(True)
    out = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, Nc, Nf, u_strat=u_s.to(dev), u_fine=u_f.to(dev),
                        precision="fp32")
    loss = ((out["rgb"] - tgt.to(dev)) ** 2).mean() + ((out["rgb0"] - tgt.to(dev)) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5
    for grads, net in ((gc, model.coarse), (gf, model.fine)):
        ref = _flat_grads(F, grads)
        err = _rel_err(net.flat.grad.cpu(), ref)
        assert err <= 2e-3, err


def test_trainer_step_matches_oracle_adam(F, cuda_device):
    """Two single-process training steps vs the oracle's loss/grads + Adam (bf16 forward, fp32 backward)."""
    from fashion_nerf_b200.train import Trainer
    dev = cuda_device
    pc, pf = O.init_params(0), O.init_params(1)
    for p in (pc, pf):
        p["alpha_linear.bias"] += 0.1
    o, d = O.pinhole_rays(16, 16)
    R, Nc, Nf = o.shape[0], 32, 32
    g = torch.Generator().manual_seed(0)
    u_s, u_f = torch.rand(R, Nc, generator=g), torch.rand(R, Nf, generator=g)
    tgt = torch.rand(R, 3, generator=torch.Generator().manual_seed(100))
    model = F.NerfModel(F.NerfNetwork.from_state_dict(pc, dev), F.NerfNetwork.from_state_dict(pf, dev))
    tr = Trainer(model)
    st_c, st_f = {}, {}
    losses = []
    for _ in range(2):
        loss_ref, gc, gf = O.loss_and_grads(pc, pf, o, d, 2.0, 6.0, Nc, Nf, tgt, u_strat=u_s, u_fine=u_f)
        pc, pf = O.adam_step(pc, gc, st_c), O.adam_step(pf, gf, st_f)
        res = tr.step(o.to(dev), d.to(dev), tgt.to(dev), 2.0, 6.0, Nc, Nf, u_strat=u_s.to(dev), u_fine=u_f.to(dev),
                      precision="fp32")
        losses.append((res["loss"].item(), loss_ref.item()))
        assert abs(res["loss"].item() - loss_ref.item()) <= 1e-4
    ref_c = F.flatten_state_dict(pc)
    # Adam normalises the step to ~lr per parameter, so compare the parameter DELTAS coarsely and the
    # parameters tightly
    assert (model.coarse.flat.cpu() - ref_c).abs().max() <= 2.5e-4
    assert losses[1][0] < losses[0][0]                                # the step reduces the loss
