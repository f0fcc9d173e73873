"""alpha compositing (A.5) fused into the last epilogue of the bf16 network-query kernel.

The fused kernel must (a) produce the very bits of `raw` the plain kernel produces, (b) composite them as the oracle's
raw2outputs does (<= 1e-5 abs, the bar of the stand-alone compositing kernel) -- in fact with the stand-alone kernel's
own bits, whatever the alignment of rays to 128-sample tiles (both walk a ray in 32-sample blocks with the arithmetic
of csrc/composite_math.cuh) -- and (c) leave the maps untouched when the raw tap is switched off (raw never reaches
HBM then)."""
import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F(cuda_device):
    import fashion_nerf_b200 as f
    f.load_library()
    return f


def _case(seed, R, S):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(R, 3, generator=g) * 0.3
    d = torch.randn(R, 3, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    nz = torch.randn(R, S, generator=g)
    return o, d, z, nz


def test_supported_sample_counts(F, cuda_device):
    ok = [s for s in range(1, 2200) if F.ops.mlp_fwd_composite_supported(s)]
    want = sorted(set(range(32, 513, 32)) | set(range(64, 1025, 64)) | set(range(128, 2049, 128)))   # S % 32 == 0, <= 16 tiles
    assert ok == want


# R chosen so that the last group of tiles is short, the last tile ragged, and CTAs own several groups
@pytest.mark.parametrize("R,S,white,noise", [(4099, 64, False, False), (2731, 192, True, False), (1237, 32, False, True),
                                             (1001, 96, True, True), (515, 128, False, False), (777, 160, False, True),
                                             (301, 256, True, False), (97, 1024, False, False), (1, 64, False, False),
                                             (3, 192, True, True)])
def test_fused_kernel_vs_plain_kernel_and_oracle(F, cuda_device, R, S, white, noise):
    dev = cuda_device
    o, d, z, nz = _case(S + R, R, S)
    nz = nz if noise else None
    net = F.NerfNetwork.random(7, dev)
    vd, dn = F.ops.ray_setup(d.to(dev))
    args = (net.packed, o.to(dev), d.to(dev), vd)
    raw = F.ops.mlp_fwd(*args, z.to(dev), precision="bf16")
    sep = F.ops.composite_fwd(raw, z.to(dev), dn, white_bkgd=white, raw_noise=None if nz is None else nz.to(dev))
    fus = F.ops.mlp_fwd_composite(*args, dn, z.to(dev), white_bkgd=white, raw_noise=None if nz is None else nz.to(dev),
                                  want_raw=True)
    torch.cuda.synchronize()
    assert torch.equal(fus["raw"], raw)                                   # (a) same bits as the plain query
    ref = O.raw2outputs(raw.cpu(), z, dn.cpu(), white, nz)                # (b) the oracle's A.5 on those bits
    for k in ("rgb", "acc", "weights"):
        assert (fus[k].cpu() - ref[k]).abs().max() <= 1e-5, k
    assert (fus["depth"].cpu() - ref["depth"]).abs().max() <= 6e-5        # depth = sum w*z, z <= 6
    for k in ("rgb", "acc", "weights", "depth"):                          # the stand-alone kernel's bits
        assert torch.equal(fus[k], sep[k]), k
    assert torch.equal(fus["disp"].nan_to_num(-1.0), sep["disp"].nan_to_num(-1.0))
    solid = ref["acc"] > 1e-3                                             # disp = acc / depth is ill-conditioned below
    assert torch.allclose(fus["disp"].cpu()[solid], ref["disp"][solid], rtol=1e-4, atol=1e-6)
    # (c) without the taps: identical maps, nothing else allocated
    lean = F.ops.mlp_fwd_composite(*args, dn, z.to(dev), white_bkgd=white, raw_noise=None if nz is None else nz.to(dev),
                                   want_raw=False, want_weights=False)
    assert lean["raw"] is None and lean["weights"] is None
    for k in ("rgb", "acc", "depth"):
        assert torch.equal(lean[k], fus[k]), k
    assert torch.equal(lean["disp"].nan_to_num(-1.0), fus["disp"].nan_to_num(-1.0))


def test_fused_kernel_many_groups_per_cta(F, cuda_device):
    """More whole-ray groups than CTAs (148): every CTA walks several groups; carry and ray sums must reset between them."""
    dev = cuda_device
    R, S = 148 * 2 * 5 + 13, 192
    o, d, z, _ = _case(3, R, S)
    net = F.NerfNetwork.random(8, dev)
    vd, dn = F.ops.ray_setup(d.to(dev))
    raw = F.ops.mlp_fwd(net.packed, o.to(dev), d.to(dev), vd, z.to(dev), precision="bf16")
    sep = F.ops.composite_fwd(raw, z.to(dev), dn)
    fus = F.ops.mlp_fwd_composite(net.packed, o.to(dev), d.to(dev), vd, dn, z.to(dev), want_raw=True)
    assert torch.equal(fus["raw"], raw)
    for k in ("rgb", "acc", "weights", "depth"):
        assert torch.equal(fus[k], sep[k]), k
    # determinism: a second launch gives the same bits
    again = F.ops.mlp_fwd_composite(net.packed, o.to(dev), d.to(dev), vd, dn, z.to(dev), want_raw=False)
    assert torch.equal(again["rgb"], fus["rgb"]) and torch.equal(again["weights"], fus["weights"])


def test_fused_kernel_empty_space_and_opaque_far_sample(F, cuda_device):
    """sigma <= 0 everywhere -> zero weights, NaN disparity (0/0) like the oracle; noise that makes only the far sample opaque."""
    dev = cuda_device
    R, S = 130, 64
    o, d, z, _ = _case(4, R, S)
    net = F.NerfNetwork.random(9, dev)
    vd, dn = F.ops.ray_setup(d.to(dev))
    raw = F.ops.mlp_fwd(net.packed, o.to(dev), d.to(dev), vd, z.to(dev), precision="bf16")
    nz = (-raw[..., 3] - 1.0).contiguous()                  # sigma + noise = -1 everywhere
    fus = F.ops.mlp_fwd_composite(net.packed, o.to(dev), d.to(dev), vd, dn, z.to(dev), raw_noise=nz, white_bkgd=True)
    assert torch.equal(fus["weights"], torch.zeros(R, S, device=dev))
    assert torch.equal(fus["rgb"], torch.ones(R, 3, device=dev)) and torch.isnan(fus["disp"]).all()
    nz[:, -1] += 2.0                                        # far sample: sigma = +1, dist = 1e10 -> alpha = 1
    fus = F.ops.mlp_fwd_composite(net.packed, o.to(dev), d.to(dev), vd, dn, z.to(dev), raw_noise=nz)
    ref = O.raw2outputs(raw.cpu(), z, dn.cpu(), False, nz.cpu())
    assert (fus["acc"].cpu() - 1).abs().max() <= 1e-6
    assert (fus["rgb"].cpu() - ref["rgb"]).abs().max() <= 1e-5


def test_unserved_sample_count_is_rejected_and_render_rays_falls_back(F, cuda_device):
    dev = cuda_device
    R, S = 64, 67
    o, d, z, _ = _case(5, R, S)
    net = F.NerfNetwork.random(10, dev)
    vd, dn = F.ops.ray_setup(d.to(dev))
    with pytest.raises(F.FnerfError, match="-2"):
        F.ops.mlp_fwd_composite(net.packed, o.to(dev), d.to(dev), vd, dn, z.to(dev))
    model = F.NerfModel.random(dev)
    with torch.no_grad():      # Nc = 33 is not served, Nc + Nf = 96 is: the coarse pass runs the separate kernels
        a = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, 33, 63, fuse_composite=True, return_taps=True)
        b = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, 33, 63, fuse_composite=False, return_taps=True)
    assert torch.equal(a["raw_c"], b["raw_c"]) and torch.equal(a["rgb0"], b["rgb0"]) and torch.equal(a["z_f"], b["z_f"])
    assert torch.equal(a["raw_f"], b["raw_f"]) and torch.equal(a["rgb"], b["rgb"])


def test_render_rays_fused_vs_separate_and_oracle(F, cuda_device):
    """The public call: fused (the inference default) vs separate kernels, with and without taps, against the oracle."""
    dev = cuda_device
    pc, pf = O.init_params(0), O.init_params(1)
    o, d = O.pinhole_rays(48, 48)
    R, Nc, Nf = o.shape[0], 64, 128
    g = torch.Generator().manual_seed(0)
    u_s, u_f = torch.rand(R, Nc, generator=g), torch.rand(R, Nf, generator=g)
    with torch.no_grad():
        ref = O.render_rays(pc, pf, o, d, 2.0, 6.0, Nc, Nf, u_strat=u_s, u_fine=u_f, return_extras=True)
    model = F.NerfModel(F.NerfNetwork.from_state_dict(pc, dev), F.NerfNetwork.from_state_dict(pf, dev))
    kw = dict(u_strat=u_s.to(dev), u_fine=u_f.to(dev))
    with torch.no_grad():
        sep = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, Nc, Nf, fuse_composite=False, return_taps=True, **kw)
        fus = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, Nc, Nf, fuse_composite=True, return_taps=True, **kw)
        lean = F.render_rays(model, o.to(dev), d.to(dev), 2.0, 6.0, Nc, Nf, **kw)          # default: fused, no raw anywhere
    assert set(lean) == {"rgb", "disp", "acc", "depth", "rgb0", "disp0", "acc0", "z_std"}
    # the fused kernels return the separate kernels' bits at every stage, so the whole render does
    for k in ("raw_c", "rgb0", "acc0", "disp0", "depth0", "z_f", "raw_f", "rgb", "acc", "depth", "z_std"):
        assert torch.equal(fus[k].nan_to_num(-1.0), sep[k].nan_to_num(-1.0)), k
    for k in ("rgb", "acc", "depth", "rgb0", "acc0", "z_std"):
        assert torch.equal(lean[k], fus[k]), k
    # against the oracle, at the bf16 bar of north_star (far-sample opacity flips excluded and counted)
    so, sk = ref["extras"]["raw_f"][:, -1, 3], fus["raw_f"][:, -1, 3].cpu()
    so0, sk0 = ref["extras"]["raw_c"][:, -1, 3], fus["raw_c"][:, -1, 3].cpu()
    flip = ((sk > 0) != (so > 0)) | ((sk0 > 0) != (so0 > 0))
    assert int(flip.sum()) <= 12
    assert (lean["rgb"].cpu()[~flip] - ref["rgb"][~flip]).abs().max() <= 2e-3
    assert (lean["rgb0"].cpu()[~flip] - ref["rgb0"][~flip]).abs().max() <= 2e-3


def test_render_rays_fused_conditioned_and_backward_guard(F, cuda_device):
    dev = cuda_device
    model = F.NerfModel.random(dev, cond=True)
    o, d = (t.to(dev) for t in O.pinhole_rays(24, 24))
    R = o.shape[0]
    g = torch.Generator().manual_seed(2)
    codes = torch.randn(3, 256, generator=g).to(dev) * 0.1
    vid = torch.randint(0, 3, (R,), generator=g).to(dev)
    with torch.no_grad():
        sep = F.render_rays(model, o, d, 2.0, 6.0, 64, 128, codes, view_id=vid, fuse_composite=False, return_taps=True)
        fus = F.render_rays(model, o, d, 2.0, 6.0, 64, 128, codes, view_id=vid, fuse_composite=True, return_taps=True)
    for k in ("raw_c", "rgb0", "z_f", "raw_f", "rgb", "acc", "depth"):
        assert torch.equal(fus[k], sep[k]), k


def test_training_keeps_raw_and_explicit_fusion_with_a_tape_is_refused(F, cuda_device):
    """A.6 needs raw: with gradients recorded the default keeps the raw tensors (taped forward, separate compositing)."""
    dev = cuda_device
    model = F.NerfModel.random(dev)
    o, d = (t.to(dev) for t in O.pinhole_rays(24, 24))
    model.coarse.flat.requires_grad_(True)
    model.fine.flat.requires_grad_(True)
    out = F.render_rays(model, o, d, 2.0, 6.0, 64, 128)
    assert out["acc"].max() > 0 and out["acc0"].max() > 0, "degenerate case: empty space everywhere"
    (out["rgb"].square().mean() + out["rgb0"].square().mean()).backward()
    for net in (model.coarse, model.fine):
        assert torch.isfinite(net.flat.grad).all() and net.flat.grad.abs().max() > 0
    with pytest.raises(ValueError):
        F.render_rays(model, o, d, 2.0, 6.0, 64, 128, fuse_composite=True, save_tape=True)
    # fused forward with the raw taps kept (no tape): the recomputing backward still works and matches the taped one
    g_tape = [n.flat.grad.clone() for n in (model.coarse, model.fine)]
    for n in (model.coarse, model.fine):
        n.flat.grad = None
    out = F.render_rays(model, o, d, 2.0, 6.0, 64, 128, fuse_composite=True, save_tape=False)
    (out["rgb"].square().mean() + out["rgb0"].square().mean()).backward()
    for n, gt in zip((model.coarse, model.fine), g_tape):
        assert (n.flat.grad - gt).norm() <= 1e-4 * gt.norm()
