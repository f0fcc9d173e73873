"""Property tests of the CPU oracle (hypothesis): invariants the domain offers independent of any implementation --
the GPU kernels are held to the same oracle, so these pin the arbiter itself."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import nerf_oracle as O

_settings = settings(max_examples=25, deadline=None)


@_settings
@given(R=st.integers(1, 40), N=st.integers(2, 96), seed=st.integers(0, 2**31 - 1), lindisp=st.booleans())
def test_stratified_is_monotone_and_in_range(R, N, seed, lindisp):
    g = torch.Generator().manual_seed(seed)
    near = 0.5 + torch.rand(R, generator=g)
    far = near + 0.5 + torch.rand(R, generator=g) * 4
    z = O.stratified(near, far, torch.linspace(0, 1, N), torch.rand(R, N, generator=g), lindisp)
    assert (z[:, 1:] >= z[:, :-1]).all()
    assert (z >= near[:, None] * (1 - 1e-6)).all() and (z <= far[:, None] * (1 + 1e-6)).all()


@_settings
@given(R=st.integers(1, 30), Nc=st.integers(3, 64), Nf=st.integers(1, 96), seed=st.integers(0, 2**31 - 1))
def test_sample_pdf_invariants(R, Nc, Nf, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(2 + 4 * torch.rand(R, Nc, generator=g), -1)[0]
    w = torch.rand(R, Nc, generator=g)
    u = torch.rand(R, Nf, generator=g)
    out = O.sample_pdf(z, w, u)
    bins = 0.5 * (z[:, 1:] + z[:, :-1])
    zs = out["z_samples"]
    assert (zs >= bins[:, :1]).all() and (zs <= bins[:, -1:]).all()          # samples stay inside the mid-point bins
    assert (out["inds"] >= 0).all() and (out["inds"] <= Nc - 1).all()
    zf = out["z_f"]
    assert zf.shape == (R, Nc + Nf) and (zf[:, 1:] >= zf[:, :-1]).all()      # merged depths sorted
    assert torch.equal(torch.sort(torch.cat([z, zs], -1), -1)[0], zf)        # and a permutation of coarse + new samples
    # monotone in u: a larger u never gives a smaller depth (inverse CDF)
    us, order = torch.sort(u, -1)
    zs_sorted = torch.gather(zs, 1, order)
    assert (zs_sorted[:, 1:] >= zs_sorted[:, :-1] - 1e-6).all()


@_settings
@given(R=st.integers(1, 30), S=st.integers(1, 128), seed=st.integers(0, 2**31 - 1), white=st.booleans())
def test_compositing_invariants(R, S, seed, white):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(R, S, 4, generator=g)
    z = torch.sort(2 + 4 * torch.rand(R, S, generator=g), -1)[0]
    dn = 0.5 + torch.rand(R, generator=g)
    out = O.raw2outputs(raw, z, dn, white)
    w = out["weights"]
    assert (w >= 0).all() and (out["acc"] <= 1 + 1e-5).all()                  # weights are a sub-probability vector
    assert (out["rgb"] >= -1e-6).all() and (out["rgb"] <= 1 + 1e-5).all()
    # transmittance telescopes: acc = 1 - prod(1 - alpha) up to the 1e-10 floor
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1) * dn[:, None]
    alpha = 1 - torch.exp(-torch.relu(raw[..., 3]) * dists)
    assert torch.allclose(out["acc"], 1 - torch.prod(1 - alpha, -1), atol=2e-5)
    # empty space renders nothing
    raw0 = raw.clone()
    raw0[..., 3] = -1.0
    o0 = O.raw2outputs(raw0, z, dn, white)
    assert (o0["acc"] == 0).all() and torch.equal(o0["rgb"], torch.full_like(o0["rgb"], 1.0 if white else 0.0))
