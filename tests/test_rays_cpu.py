"""The caller side of render_rays: the package's synthetic pinhole camera must produce the oracle's
rays bit for bit (the two functions are deliberate twins: the product may not import the oracle)."""
import torch

import fashion_nerf_b200 as F
from oracle import nerf_oracle as O


def test_pinhole_rays_equal_the_oracles_bit_for_bit():
    for (H, W, view, n_views) in ((64, 64, 0, 1), (8, 16, 1, 4), (800, 800, 0, 1), (24, 40, 7, 8), (512, 512, 31, 32)):
        o, d = F.pinhole_rays(H, W, view=view, n_views=n_views)
        ro, rd = O.pinhole_rays(H, W, view=view, n_views=n_views)
        assert o.shape == (H * W, 3) and d.dtype == torch.float32
        assert torch.equal(o, ro) and torch.equal(d, rd), (H, W, view, n_views)


def test_pinhole_rays_geometry():
    """Un-normalised directions (|d| > 1 off-axis so the dnorm factor of A.5 is exercised), origin at distance 4,
    views orbit about +y."""
    o, d = F.pinhole_rays(64, 64)
    assert torch.equal(o, torch.tensor([0.0, 0.0, 4.0]).expand(4096, 3))
    assert (d[:, 2] == -1).all() and d.norm(dim=-1).max() > 1.05
    o2, d2 = F.pinhole_rays(64, 64, view=2, n_views=8)            # 90 degrees about +y
    assert torch.allclose(o2, torch.tensor([4.0, 0.0, 0.0]).expand(4096, 3), atol=1e-6)
    assert torch.allclose(o2.norm(dim=-1), torch.full((4096,), 4.0), atol=1e-6)
    assert torch.allclose(d2.norm(dim=-1), d.norm(dim=-1), atol=1e-6)
