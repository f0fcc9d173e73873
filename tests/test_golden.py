"""Golden fixtures (tests/golden/*.pt, written by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them (bit-exact where the contract is bit-exact).  GPU: the CUDA
kernels, called through the C ABI, reproduce them too."""
import os

import pytest
import torch

from oracle import nerf_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sc():
    return torch.load(os.path.join(G, "sampling_compositing.pt"))


@pytest.fixture(scope="module")
def nq():
    return torch.load(os.path.join(G, "network_query.pt"))


def test_oracle_reproduces_golden_sampling(sc):
    z = O.stratified(sc["near"], sc["far"], sc["t_vals"], sc["u_strat"])
    assert torch.equal(z, sc["z"])
    sp = O.sample_pdf(sc["z"], sc["weights"], sc["u_fine"])
    assert torch.equal(sp["inds"].int(), sc["inds"])
    assert torch.equal(sp["z_samples"], sc["z_samples"]) and torch.equal(sp["z_f"], sc["z_f"])


def test_oracle_reproduces_golden_compositing(sc):
    out = O.raw2outputs(sc["raw"], sc["z"], sc["dnorm"])
    for k in ("rgb", "depth", "acc", "weights"):
        assert (out[k] - sc[k]).abs().max() <= 2e-6, k
    g = O.composite_bwd(sc["raw"], sc["z"], sc["dnorm"], sc["g_rgb"], sc["g_depth"], sc["g_acc"])
    assert (g - sc["g_raw_fp64"]).abs().max() <= 1e-4 * max(1.0, sc["g_raw_fp64"].abs().max().item())


def test_oracle_reproduces_golden_network(nq):
    p = O.init_params(nq["seed"])
    pts = nq["rays_o"][:, None, :] + nq["rays_d"][:, None, :] * nq["z"][:, :, None]
    raw = O.run_network(p, pts, nq["viewdirs"])
    assert (raw - nq["raw_fp64"]).abs().max() <= 1e-5


@pytest.mark.gpu
def test_kernels_reproduce_golden(sc, nq, cuda_device):
    import fashion_nerf_b200 as F
    dev = cuda_device
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
    z = F.ops.stratified(d["near"], d["far"], d["t_vals"], d["u_strat"])
    assert torch.equal(z.cpu(), sc["z"])
    sp = F.ops.importance(d["z"], d["weights"], d["u_fine"])
    assert torch.equal(sp["inds"].cpu(), sc["inds"])
    assert torch.equal(sp["z_samples"].cpu(), sc["z_samples"]) and torch.equal(sp["z_f"].cpu(), sc["z_f"])
    out = F.ops.composite_fwd(d["raw"], d["z"], d["dnorm"])
    for k in ("rgb", "acc", "weights"):
        assert (out[k].cpu() - sc[k]).abs().max() <= 1e-5, k
    g = F.ops.composite_bwd(d["raw"], d["z"], d["dnorm"], d["g_rgb"], d["g_depth"], d["g_acc"]).cpu()
    assert (g - sc["g_raw_fp64"]).abs().max() <= 1e-4 * max(1.0, sc["g_raw_fp64"].abs().max().item())
    net = F.NerfNetwork.random(nq["seed"], dev)
    q = [nq[k].to(dev) for k in ("rays_o", "rays_d", "viewdirs", "z")]
    assert (F.ops.mlp_fwd(net.packed, *q, precision="fp32").cpu() - nq["raw_fp64"]).abs().max() <= 2e-5
    assert (F.ops.mlp_fwd(net.packed, *q, precision="bf16").cpu() - nq["raw_fp64"]).abs().max() <= 2e-3
