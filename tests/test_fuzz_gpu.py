"""Shape / value fuzzing of the bit-exact stages against the oracle (SURVEY.md 4 "Edge cases": ragged N and R, u at the
ends of [0,1), zero and spiky weights, near == far).  hypothesis drives the shapes; every example is tiny."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
_settings = settings(max_examples=30, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@pytest.fixture(scope="module")
def F(cuda_device):
    import fashion_nerf_b200 as f
    f.load_library()
    return f


@_settings
@given(R=st.integers(1, 300), N=st.integers(1, 200), jitter=st.booleans(), lindisp=st.booleans(), seed=st.integers(0, 2**31 - 1),
       degenerate=st.booleans())
def test_fuzz_stratified(F, cuda_device, R, N, jitter, lindisp, seed, degenerate):
    g = torch.Generator().manual_seed(seed)
    near = 0.5 + torch.rand(R, generator=g) * 2
    far = near.clone() if degenerate else near + 0.1 + torch.rand(R, generator=g) * 5
    t = torch.linspace(0, 1, N)
    u = torch.rand(R, N, generator=g) if jitter else None
    if u is not None and R * N > 2:
        u.view(-1)[0] = 0.0
        u.view(-1)[-1] = 1.0 - 2.0 ** -24
    ref = O.stratified(near, far, t, u, lindisp)
    dev = cuda_device
    got = F.ops.stratified(near.to(dev), far.to(dev), t.to(dev), None if u is None else u.to(dev), lindisp)
    assert torch.equal(got.cpu(), ref)


@_settings
@given(R=st.integers(1, 120), Nc=st.integers(3, 96), Nf=st.integers(1, 160), seed=st.integers(0, 2**31 - 1),
       wkind=st.sampled_from(["rand", "zero", "spike", "tiny"]))
def test_fuzz_importance(F, cuda_device, R, Nc, Nf, seed, wkind):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(2 + 4 * torch.rand(R, Nc, generator=g), -1)[0]
    if wkind == "rand":
        w = torch.rand(R, Nc, generator=g)
    elif wkind == "zero":
        w = torch.zeros(R, Nc)
    elif wkind == "tiny":
        w = torch.rand(R, Nc, generator=g) * 1e-7
    else:
        w = torch.zeros(R, Nc)
        w[torch.arange(R), torch.randint(0, Nc, (R,), generator=g)] = 1.0
    u = torch.rand(R, Nf, generator=g)
    u.view(-1)[0] = 0.0
    u.view(-1)[-1] = 1.0 - 2.0 ** -24
    ref = O.sample_pdf(z, w, u)
    dev = cuda_device
    out = F.ops.importance(z.to(dev), w.to(dev), u.to(dev))
    assert torch.equal(out["z_samples"].cpu(), ref["z_samples"])
    assert torch.equal(out["inds"].cpu().long(), ref["inds"])
    assert torch.equal(out["z_f"].cpu(), ref["z_f"])


@_settings
@given(R=st.integers(1, 200), S=st.integers(1, 300), white=st.booleans(), seed=st.integers(0, 2**31 - 1),
       sig=st.sampled_from(["rand", "neg", "huge"]))
def test_fuzz_composite(F, cuda_device, R, S, white, seed, sig):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(R, S, 4, generator=g)
    if sig == "neg":
        raw[..., 3] = -raw[..., 3].abs()              # sigma <= 0 everywhere: empty space
    elif sig == "huge":
        raw[..., 3] = raw[..., 3].abs() * 50.0        # opaque after the first sample
    z = torch.sort(2 + 4 * torch.rand(R, S, generator=g), -1)[0]
    dn = 0.5 + 2 * torch.rand(R, generator=g)
    ref = O.raw2outputs(raw, z, dn, white)
    dev = cuda_device
    out = F.ops.composite_fwd(raw.to(dev), z.to(dev), dn.to(dev), white_bkgd=white)
    for k in ("rgb", "acc", "weights"):
        assert (out[k].cpu() - ref[k]).abs().max() <= 1e-5, k
    assert (out["depth"].cpu() - ref["depth"]).abs().max() <= 6e-5   # depth sums w * z with z up to 6
    g_rgb, g_d, g_a = torch.randn(R, 3, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g)
    ref64 = O.composite_bwd(raw.double(), z.double(), dn.double(), g_rgb.double(), g_d.double(), g_a.double(), white)
    ref32 = O.composite_bwd(raw, z, dn, g_rgb, g_d, g_a, white)
    got = F.ops.composite_bwd(raw.to(dev), z.to(dev), dn.to(dev), g_rgb.to(dev), g_d.to(dev), g_a.to(dev), white_bkgd=white).cpu()
    assert torch.isfinite(got).all()
    err_kernel = (got.double() - ref64).abs().max().item()
    err_oracle32 = (ref32.double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    # as close to the fp64 truth as the fp32 oracle is (the far sample's 1e10 distance makes the problem ill-conditioned)
    assert err_kernel <= max(4 * err_oracle32, 2e-5 * max(scale, 1.0)), (err_kernel, err_oracle32, scale)
