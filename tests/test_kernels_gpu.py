"""Stage-isolated parity of every C-ABI entry against the CPU oracle on identical inputs
("Kernel unit parity").  All calls go through libfnerf.so via ctypes.

Bars (BASELINE.json north_star): bit-exact sample positions and importance-bin indices;
compositing <= 1e-5 abs in fp32; fp32 MLP <= 2e-5 abs on raw; bf16 MLP checked end to end in
test_render_gpu.py."""
import math

import pytest
import torch

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F(cuda_device):
    import fashion_nerf_b200 as f
    f.load_library()
    return f


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------ A.1
def test_ray_setup_bit_exact(F, cuda_device):
    _, d = O.pinhole_rays(37, 53)
    d = torch.cat([d, torch.randn(1000, 3, generator=_gen(0)) * 3])
    vd_ref, dn_ref = O.ray_setup_exact(d)
    vd, dn = F.ops.ray_setup(d.to(cuda_device))
    assert torch.equal(vd.cpu(), vd_ref) and torch.equal(dn.cpu(), dn_ref)


# ------------------------------------------------------------------------------------------ A.2
@pytest.mark.parametrize("R,N,jitter,lindisp", [(4096, 64, True, False), (4096, 64, False, False),
                                                (1001, 256, True, False), (333, 7, True, True),
                                                (5, 1, True, False), (64, 64, False, True), (100, 32, True, True),
                                                (77, 12, True, False), (300000, 64, True, False)])
def test_stratified_bit_exact(F, cuda_device, R, N, jitter, lindisp):
    g = _gen(1)
    near = 1.5 + torch.rand(R, generator=g)
    far = 5.0 + torch.rand(R, generator=g) * 2
    t = torch.linspace(0, 1, N)
    u = torch.rand(R, N, generator=g) if jitter else None
    ref = O.stratified(near, far, t, u, lindisp)
    got = F.ops.stratified(near.to(cuda_device), far.to(cuda_device), t.to(cuda_device),
                           None if u is None else u.to(cuda_device), lindisp)
    assert torch.equal(got.cpu(), ref)


def test_stratified_near_equals_far(F, cuda_device):
    near = torch.full((8,), 3.0)
    t = torch.linspace(0, 1, 16)
    u = torch.rand(8, 16, generator=_gen(2))
    ref = O.stratified(near, near, t, u)
    got = F.ops.stratified(near.to(cuda_device), near.to(cuda_device), t.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got.cpu(), ref)


# ------------------------------------------------------------------------------------------ A.7
def _coarse_case(R, Nc, seed, peaky=True):
    g = _gen(seed)
    near, far = torch.full((R,), 2.0), torch.full((R,), 6.0)
    z = O.stratified(near, far, torch.linspace(0, 1, Nc), torch.rand(R, Nc, generator=g))
    raw = torch.randn(R, Nc, 4, generator=g) * (4.0 if peaky else 1.0)
    w = O.raw2outputs(raw, z, torch.ones(R))["weights"]
    return z, w


@pytest.mark.parametrize("R,Nc,Nf", [(4096, 64, 128), (513, 256, 768), (100, 17, 5), (64, 3, 1), (7, 64, 200)])
def test_importance_bit_exact(F, cuda_device, R, Nc, Nf):
    z, w = _coarse_case(R, Nc, 3)
    u = torch.rand(R, Nf, generator=_gen(4))
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])
    assert torch.allclose(got["z_std"].cpu(), ref["z_std"], rtol=1e-4, atol=1e-6)


def test_importance_shared_linspace_row(F, cuda_device):
    z, w = _coarse_case(777, 64, 5)
    u1 = torch.linspace(0, 1, 128)
    ref = O.sample_pdf(z, w, u1[None].expand(777, 128))
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u1.to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])


def test_importance_edge_cases(F, cuda_device):
    """u in {0, just below 1, 1, > cdf[-1]}, all-zero weights, one-hot weights, constant z."""
    Nc = 64
    z = torch.linspace(2, 6, Nc)[None].repeat(4, 1).contiguous()
    z[3] = 4.0                                                    # near == far
    w = torch.zeros(4, Nc)
    w[1, 30] = 1.0
    w[2] = torch.rand(Nc, generator=_gen(6)) * 1e-7
    u = torch.tensor([0.0, 1e-8, 0.25, 0.5, 0.99999994, 1.0, 1.0000001, 1.5])[None].repeat(4, 1).contiguous()
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])


def test_importance_full_size_sortedness(F, cuda_device):
    """Size-independent property at the long-ray size (cfg 4): output sorted, a permutation of inputs."""
    R, Nc, Nf = 20000, 256, 768
    g = torch.Generator(device="cpu").manual_seed(7)
    z = torch.sort(torch.rand(R, Nc, generator=g) * 4 + 2, -1)[0].to(cuda_device)
    w = torch.rand(R, Nc, generator=g).to(cuda_device)
    u = torch.rand(R, Nf, generator=g).to(cuda_device)
    got = F.ops.importance(z, w, u)
    zf = got["z_f"]
    assert (zf[:, 1:] >= zf[:, :-1]).all()
    assert torch.equal(zf, torch.sort(torch.cat([z, got["z_samples"]], -1), -1)[0])
    assert (got["inds"] >= 1).all() and (got["inds"] <= Nc - 1).all()


# ------------------------------------------------------------------------------------------ A.3
def test_posenc(F, cuda_device):
    x = torch.rand(5000, 3, generator=_gen(8)) * 12 - 6
    for L in (10, 4, 0):
        ref = O.posenc(x, L) if L > 0 else x
        got = F.ops.posenc(x.to(cuda_device), L).cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max() <= 1e-6


# ------------------------------------------------------------------------------------------ A.5
@pytest.mark.parametrize("R,S,white,noise", [(4096, 64, False, False), (4096, 192, False, False),
                                             (257, 1024, True, False), (1000, 33, True, True), (3, 1, False, False)])
def test_composite_fwd(F, cuda_device, R, S, white, noise):
    g = _gen(9)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, generator=g)
    dn = 1 + torch.rand(R, generator=g)
    nz = torch.randn(R, S, generator=g) if noise else None
    ref = O.raw2outputs(raw, z, dn, white, nz)
    got = F.ops.composite_fwd(raw.to(cuda_device), z.to(cuda_device), dn.to(cuda_device), white_bkgd=white,
                              raw_noise=None if nz is None else nz.to(cuda_device))
    for k in ("rgb", "acc", "weights"):
        assert (got[k].cpu() - ref[k]).abs().max() <= 1e-5, k
    assert (got["depth"].cpu() - ref["depth"]).abs().max() <= 1e-5 * 6       # depth = sum w*z, z <= 6
    assert torch.allclose(got["disp"].cpu(), ref["disp"], rtol=1e-4, atol=1e-6, equal_nan=True)


def test_composite_fwd_empty_space(F, cuda_device):
    R, S = 64, 64
    z = torch.linspace(2, 6, S)[None].repeat(R, 1).contiguous()
    raw = -torch.rand(R, S, 4, generator=_gen(10))                    # sigma <= 0 everywhere
    ref = O.raw2outputs(raw, z, torch.ones(R), True)
    got = F.ops.composite_fwd(raw.to(cuda_device), z.to(cuda_device), torch.ones(R, device=cuda_device), white_bkgd=True)
    assert torch.equal(got["weights"].cpu(), torch.zeros(R, S))
    assert torch.equal(got["rgb"].cpu(), ref["rgb"])
    assert torch.isnan(got["disp"]).all() and torch.isnan(ref["disp"]).all()


def test_composite_fwd_constant_sigma_closed_form(F, cuda_device):
    """Analytic check independent of the oracle: acc = 1, rgb = sigmoid(c) for constant raw."""
    R, S = 128, 192
    z = torch.linspace(2, 6, S)[None].repeat(R, 1).contiguous()
    raw = torch.zeros(R, S, 4)
    raw[..., 3] = 0.7
    raw[..., :3] = torch.tensor([0.3, -0.2, 1.1])
    got = F.ops.composite_fwd(raw.to(cuda_device), z.to(cuda_device), torch.ones(R, device=cuda_device))
    assert (got["acc"].cpu() - 1).abs().max() < 1e-5
    assert (got["rgb"].cpu() - torch.sigmoid(torch.tensor([0.3, -0.2, 1.1]))).abs().max() < 1e-5


# ------------------------------------------------------------------------------------------ A.6
@pytest.mark.parametrize("R,S,white", [(1024, 64, False), (512, 192, True), (65, 1024, False), (9, 33, True)])
def test_composite_bwd(F, cuda_device, R, S, white):
    g = _gen(11)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, generator=g)
    dn = 1 + torch.rand(R, generator=g)
    g_rgb, g_d, g_a = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g)
    ref64 = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double(), white)
    ref32 = O.composite_bwd(raw, z, dn, g_rgb, g_d, g_a, white)
    got = F.ops.composite_bwd(raw.to(cuda_device), z.to(cuda_device), dn.to(cuda_device), g_rgb.to(cuda_device),
                              g_d.to(cuda_device), g_a.to(cuda_device), white_bkgd=white).cpu()
    assert torch.isfinite(got).all()
    err_kernel = (got.double() - ref64).abs().max().item()
    err_oracle32 = (ref32.double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    # the fp32 kernel must be as close to the fp64 truth as the fp32 oracle is (same conditioning)
    assert err_kernel <= max(4 * err_oracle32, 1e-5 * max(scale, 1.0)), (err_kernel, err_oracle32, scale)


# ------------------------------------------------------------------------------------------ weights
@pytest.mark.parametrize("cond", [False, True])
def test_pack_unpack_round_trip(F, cuda_device, cond):
    sd = F.init_state_dict(3, cond)
    flat = F.flatten_state_dict(sd, cond)
    assert flat.numel() == F.ops.param_count(cond)
    packed = F.ops.pack_weights(flat.to(cuda_device), cond)
    back = F.ops.unpack_weights(packed, cond).cpu()
    assert torch.equal(back, flat)


def test_init_matches_oracle_init():
    import fashion_nerf_b200 as f
    a, b = f.init_state_dict(1), O.init_params(1)
    assert all(torch.equal(a[k], b[k]) for k in b)


# ------------------------------------------------------------------------------------------ A.4 fp32
def _query_case(R, S, seed):
    g = _gen(seed)
    o = torch.rand(R, 3, generator=g) * 2 - 1
    d = torch.randn(R, 3, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    vd, _ = O.ray_setup_exact(d)
    return o, d, vd, z


@pytest.mark.parametrize("R,S", [(64, 64), (33, 7), (10, 192)])
def test_mlp_fp32_vs_oracle(F, cuda_device, R, S):
    p = O.init_params(0)
    o, d, vd, z = _query_case(R, S, 12)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    ref = O.run_network(p, pts, vd)
    net = F.NerfNetwork.from_state_dict(p, cuda_device)
    got = F.ops.mlp_fwd(net.packed, o.to(cuda_device), d.to(cuda_device), vd.to(cuda_device), z.to(cuda_device),
                        precision="fp32").cpu()
    assert (got - ref).abs().max() <= 2e-5, (got - ref).abs().max()


def test_mlp_fp32_cond_vs_oracle(F, cuda_device):
    R, S, V = 48, 16, 4
    p = O.init_params(2, cond=True)
    o, d, vd, z = _query_case(R, S, 13)
    cond = torch.randn(V, 256, generator=_gen(14))
    view_id = torch.randint(0, V, (R,), generator=_gen(15))
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    ref = O.run_network(p, pts, vd, cond[view_id])
    net = F.NerfNetwork.from_state_dict(p, cuda_device, cond=True)
    proj = F.ops.cond_project(net.packed, cond.to(cuda_device))
    got = F.ops.mlp_fwd(net.packed, o.to(cuda_device), d.to(cuda_device), vd.to(cuda_device), z.to(cuda_device),
                        precision="fp32", cond_proj=proj, cond_index=view_id.to(cuda_device)).cpu()
    assert (got - ref).abs().max() <= 3e-5, (got - ref).abs().max()


# ------------------------------------------------------------------------------------------ errors
def test_error_codes_and_no_cpu_fallback(F, cuda_device):
    lib = F.load_library()
    assert lib.fnerf_stratified(None, None, None, None, None, 4, 4, 0, None) == -1
    assert b"null" in lib.fnerf_last_error()
    assert lib.fnerf_importance(None, None, None, 0, None, None, None, None, 4, 2, 4, None) == -2
    with pytest.raises(F.FnerfError):
        F.ops.ray_setup(torch.zeros(4, 3))                              # CPU tensor: rejected, not emulated


def test_error_codes_training_entries(F, cuda_device):
    """Argument validation of the ABI v2 training entries: negative codes before any launch."""
    lib = F.load_library()
    dev = cuda_device
    net = F.NerfNetwork.random(0, dev)
    R, S = 8, 16
    o = torch.zeros(R, 3, device=dev); d = torch.ones(R, 3, device=dev); z = torch.ones(R, S, device=dev)
    raw = torch.empty(R, S, 4, device=dev)
    tape = torch.empty(F.ops.mlp_tape_bytes(R, S), dtype=torch.uint8, device=dev)
    p = lambda t: t.data_ptr()
    assert F.ops.mlp_tape_bytes(R, S) == 40 * 16384 + 68 * 128 * 4          # one 128-sample tile
    # tape too small / null tape / misaligned tape
    assert lib.fnerf_mlp_fwd_tape(p(net.packed), 0, p(o), p(d), p(d), p(z), None, None, 0, p(raw), p(tape), tape.numel() - 1, R, S, None) == -5
    assert lib.fnerf_mlp_fwd_tape(p(net.packed), 0, p(o), p(d), p(d), p(z), None, None, 0, p(raw), None, tape.numel(), R, S, None) == -1
    assert lib.fnerf_mlp_fwd_tape(p(net.packed), 0, p(o), p(d), p(d), p(z), None, None, 0, p(raw), p(tape) + 4, tape.numel(), R, S, None) == -3
    g = torch.zeros(R, S, 4, device=dev); fg = torch.zeros(net.flat.numel(), device=dev)
    ws = torch.empty(int(lib.fnerf_mlp_bwd_tape_workspace_bytes(R, S)), dtype=torch.uint8, device=dev)
    assert lib.fnerf_mlp_bwd_tape(p(net.packed), 0, p(g), p(tape), tape.numel(), None, None, 0, p(fg), p(ws), ws.numel() - 1, R, S, None) == -5
    assert lib.fnerf_mlp_bwd_tape(p(net.packed), 1, p(g), p(tape), tape.numel(), None, None, 0, p(fg), p(ws), ws.numel(), R, S, None) == -1   # cond without codes
    assert lib.fnerf_mlp_bwd_tape(p(net.packed), 0, p(g), p(tape), tape.numel(), None, None, 0, p(fg), p(ws), ws.numel(), 0, S, None) == 0    # R == 0: no-op
    assert lib.fnerf_adam_step(p(fg), p(fg), p(fg), p(fg), fg.numel(), 1e-3, 0.9, 0.999, 1e-8, 0, 1.0, None) == -2                            # step must be >= 1
    assert lib.fnerf_adam_step(None, p(fg), p(fg), p(fg), fg.numel(), 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None) == -1
    with pytest.raises(ValueError):
        F.render_rays(F.NerfModel(net, None), o, d, 2.0, 6.0, 8, 0, precision="fp32", save_tape=True)
    assert fg.abs().max() == 0                                              # nothing above touched the gradient buffer


def test_importance_descending_coarse_depths(F, cuda_device):
    """near > far gives descending coarse depths: the rank-merge fast path must hand over to the
    generic sort, and odd sizes must still be bit-exact."""
    for R, Nc, Nf in ((257, 64, 128), (33, 20, 45)):
        g = _gen(31)
        near, far = torch.full((R,), 6.0), torch.full((R,), 2.0)
        z = O.stratified(near, far, torch.linspace(0, 1, Nc), torch.rand(R, Nc, generator=g))
        assert (z[:, 1:] < z[:, :-1]).all()
        w = torch.rand(R, Nc, generator=g)
        u = torch.rand(R, Nf, generator=g)
        ref = O.sample_pdf(z, w, u)
        got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
        assert torch.equal(got["inds"].cpu().long(), ref["inds"])
        assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
        assert torch.equal(got["z_f"].cpu(), ref["z_f"])


def test_importance_ties_and_duplicates(F, cuda_device):
    """Duplicate u values and samples equal to coarse depths: merged output must equal torch.sort's."""
    R, Nc, Nf = 64, 32, 96
    g = _gen(32)
    z = torch.sort(torch.rand(R, Nc, generator=g) * 4 + 2, -1)[0]
    z[:, 10] = z[:, 9]                                                 # duplicate coarse depths
    w = torch.rand(R, Nc, generator=g)
    u = torch.rand(R, Nf, generator=g)
    u[:, 1::2] = u[:, 0::2]                                            # duplicate uniforms -> duplicate samples
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])


@pytest.mark.parametrize("R,S,white", [(300, 64, False), (40, 300, True)])
def test_composite_bwd_with_raw_noise(F, cuda_device, R, S, white):
    """A.6 with the forward's raw_noise: sigma = raw[...,3] + noise decides alpha AND the ReLU gate (both the
    register-resident and the shared-memory kernel)."""
    g = torch.Generator().manual_seed(S)
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    dn = 1 + torch.rand(R, generator=g)
    nz = torch.randn(R, S, generator=g)
    g_rgb, g_d, g_a = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g)
    want = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double(), white, nz.double())
    dev = cuda_device
    got = F.ops.composite_bwd(raw.to(dev), z.to(dev), dn.to(dev), g_rgb.to(dev), g_d.to(dev), g_a.to(dev), white_bkgd=white,
                              raw_noise=nz.to(dev)).cpu()
    ref32 = O.composite_bwd(raw, z, dn, g_rgb, g_d, g_a, white, nz)
    scale = want.abs().max().item()
    assert torch.isfinite(got).all()
    assert (got.double() - want).abs().max().item() <= max(4 * (ref32.double() - want).abs().max().item(), 1e-5 * max(scale, 1.0))
    plain = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double(), white)
    assert (plain - want).abs().max().item() > 1e-2 * scale          # the noise changes the answer: the test has teeth
    fwd = F.ops.composite_fwd(raw.to(dev), z.to(dev), dn.to(dev), white_bkgd=white, raw_noise=nz.to(dev))
    ref = O.raw2outputs(raw, z, dn, white, nz)
    assert (fwd["rgb"].cpu() - ref["rgb"]).abs().max() <= 1e-5 and (fwd["weights"].cpu() - ref["weights"]).abs().max() <= 1e-5


@pytest.mark.parametrize("Nc,Nf", [(64, 128), (32, 32), (64, 64), (128, 128), (128, 256), (256, 768)])
@pytest.mark.parametrize("kind", ["random", "linspace_row", "sorted", "peaky", "descending"])
def test_importance_register_path_bit_exact(F, cuda_device, Nc, Nf, kind):
    """The register-resident kernels (Nc = 32*2^a, Nf = 32*b; 256 + 768 pads its 24 samples per lane to 32): bit-exact
    indices, samples and merged depths for random uniforms, the shared deterministic row (sort skipped), per-ray sorted
    uniforms, peaky weights (many samples in one bin, duplicates) and descending coarse depths (slow in-kernel path)."""
    R = 777 if Nf < 768 else 301
    g = _gen(Nc * 1000 + Nf)
    if kind == "descending":
        z = O.stratified(torch.full((R,), 6.0), torch.full((R,), 2.0), torch.linspace(0, 1, Nc), torch.rand(R, Nc, generator=g))
        w = torch.rand(R, Nc, generator=g)
    else:
        z, w = _coarse_case(R, Nc, Nc + Nf, peaky=(kind == "peaky"))
    if kind == "linspace_row":
        u_dev = torch.linspace(0, 1, Nf)
        u = u_dev[None].expand(R, Nf)
    else:
        u = torch.rand(R, Nf, generator=g)
        if kind == "sorted":
            u = torch.sort(u, -1)[0]
        if kind == "peaky":
            u[:, 1::2] = u[:, 0::2]
        u_dev = u
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u_dev.contiguous().to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])
    assert torch.allclose(got["z_std"].cpu(), ref["z_std"], rtol=1e-4, atol=1e-6)


def test_importance_fast_division_matches_ieee(F, cuda_device):
    """The inverse CDF divides with div.rn.f32's fast-path instruction sequence behind its own range test
    (sampling.cu fdiv_rn_inrange): bit-identical to IEEE division on 2^28 pairs of the ranges the kernel produces."""
    lib = F.load_library()
    bad = torch.zeros(2, dtype=torch.int64, device=cuda_device)
    for seed in (1, 2):
        assert lib.fnerf_debug_fdiv_mismatches(1 << 27, seed, bad.data_ptr(), None) == 0
    torch.cuda.synchronize()
    n_bad, pair = int(bad[0].item()), int(bad[1].item()) & (2 ** 64 - 1)
    assert n_bad == 0, f"{n_bad} mismatches, e.g. x bits {pair >> 32:#010x}, d bits {pair & 0xffffffff:#010x}"


def test_importance_tiny_and_huge_uniforms_take_the_ieee_division(F, cuda_device):
    """Numerators outside the fast division's range (denormal / tiny u in the first bin, u far above 1, negative u) go
    through __fdiv_rn: still bit-exact on the register path."""
    R, Nc, Nf = 64, 64, 128
    z, w = _coarse_case(R, Nc, 77)
    u = torch.rand(R, Nf, generator=_gen(78))
    u[:, 3] = 1e-42
    u[:, 40] = 3e-33
    u[:, 77] = 0.0
    u[1::2, 100] = 2.5e31
    u[0::2, 101] = -1e-35
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])


@pytest.mark.parametrize("R,S,white", [(28421, 64, False), (47371, 32, True), (80003, 192, True), (76543, 100, False)])
def test_composite_fwd_multi_ray_passes(F, cuda_device, R, S, white):
    """Launch sizes at which a warp of the register-resident forward walks SEVERAL rays per pass (lane j keeps the j-th
    ray's totals, one lane per ray runs the scalar tail): consecutive rays for S <= 64 (pass length R / (64 warps x SMs),
    ragged last pass), rays one grid apart above.  Per-ray maps against the oracle, ray by ray."""
    g = _gen(R + S)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, generator=g)
    dn = 1 + torch.rand(R, generator=g)
    ref = O.raw2outputs(raw, z, dn, white)
    got = F.ops.composite_fwd(raw.to(cuda_device), z.to(cuda_device), dn.to(cuda_device), white_bkgd=white)
    for k in ("rgb", "acc", "weights"):
        assert (got[k].cpu() - ref[k]).abs().max() <= 1e-5, k
    assert (got["depth"].cpu() - ref["depth"]).abs().max() <= 1e-5 * 6
    assert torch.allclose(got["disp"].cpu(), ref["disp"], rtol=1e-4, atol=1e-6, equal_nan=True)
    # the same rays in a small launch (one ray per pass) give the same bits
    sl = slice(R - 1000, R)
    small = F.ops.composite_fwd(raw[sl].to(cuda_device), z[sl].to(cuda_device), dn[sl].to(cuda_device), white_bkgd=white)
    for k in ("rgb", "acc", "depth", "disp", "weights"):
        assert torch.equal(small[k], got[k][sl]) or (k == "disp" and torch.equal(small[k].isnan(), got[k][sl].isnan())), k


@pytest.mark.parametrize("Nc,Nf,R", [(64, 128, 30011), (32, 32, 25007), (128, 256, 12001)])
def test_importance_register_path_many_rays_per_warp(F, cuda_device, Nc, Nf, R):
    """More rays than warps in the grid: every warp reuses its shared-memory tables (cdf, gather entries, sorted samples /
    the prefix-maximum table, merged row) for several rays in a row.  Bit-exact against the oracle, mixed sorted and
    unsorted uniform rows so that consecutive rays of a warp take the sort and the skip-sort branch."""
    z, w = _coarse_case(R, Nc, R % 97, peaky=True)
    u = torch.rand(R, Nf, generator=_gen(R))
    u[::3] = torch.sort(u[::3], -1)[0]
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(cuda_device), w.to(cuda_device), u.to(cuda_device))
    assert torch.equal(got["inds"].cpu().long(), ref["inds"])
    assert torch.equal(got["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(got["z_f"].cpu(), ref["z_f"])


@pytest.mark.parametrize("R,S,white", [(6001, 300, False), (5003, 777, True), (4801, 512, False), (4750, 1000, True)])
def test_composite_bwd_long_rays_shared_by_warps(F, cuda_device, R, S, white):
    """256 < S <= 1024: ceil(S / 128) warps share a ray and exchange the chunk transmittances / suffix sums through shared
    memory; more rays than CTAs in the grid, so every CTA reuses the exchange buffers ray after ray (parity double
    buffering).  Against the fp64 oracle, as close as the fp32 oracle is."""
    g = _gen(R + S)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, generator=g)
    dn = 1 + torch.rand(R, generator=g)
    g_rgb, g_d, g_a = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g)
    ref64 = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double(), white)
    ref32 = O.composite_bwd(raw, z, dn, g_rgb, g_d, g_a, white)
    dev = cuda_device
    got = F.ops.composite_bwd(raw.to(dev), z.to(dev), dn.to(dev), g_rgb.to(dev), g_d.to(dev), g_a.to(dev), white_bkgd=white).cpu()
    assert torch.isfinite(got).all()
    err_kernel = (got.double() - ref64).abs().max().item()
    err_oracle32 = (ref32.double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    assert err_kernel <= max(4 * err_oracle32, 1e-5 * max(scale, 1.0)), (err_kernel, err_oracle32, scale)
