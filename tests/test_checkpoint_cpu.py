"""Checkpoint dictionaries: host-side only, no GPU.  The format is the canonical nerf-pytorch one,
checked here against a from-scratch torch.nn restatement of the canonical module (same parameter names)."""
import io

import pytest
import torch

from fashion_nerf_b200 import checkpoint as C
from fashion_nerf_b200.model import flatten_state_dict, init_state_dict, unflatten


class CanonicalNeRF(torch.nn.Module):
    """NeRF(D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs=True): parameter names only."""

    def __init__(self, cond=False):
        super().__init__()
        k5 = 63 + (256 if cond else 0) + 256
        self.pts_linears = torch.nn.ModuleList([torch.nn.Linear(63, 256)] +
                                               [torch.nn.Linear(k5 if i == 4 else 256, 256) for i in range(7)])
        self.views_linears = torch.nn.ModuleList([torch.nn.Linear(256 + 27, 128)])
        self.feature_linear = torch.nn.Linear(256, 256)
        self.alpha_linear = torch.nn.Linear(256, 1)
        self.rgb_linear = torch.nn.Linear(128, 3)


@pytest.mark.parametrize("cond", [False, True])
def test_canonical_module_state_dict_round_trip(cond):
    torch.manual_seed(0)
    m = CanonicalNeRF(cond)
    sd, got_cond = C.validate_state_dict(m.state_dict())
    assert got_cond == cond
    flat = flatten_state_dict(sd, cond)
    assert flat.numel() == sum(p.numel() for p in m.parameters())
    back = unflatten(flat, cond)
    m2 = CanonicalNeRF(cond)
    m2.load_state_dict(back)                                   # names and shapes are the canonical ones
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_checkpoint_file_round_trip_and_wrappers():
    sd_c, sd_f = init_state_dict(0), init_state_dict(1)
    ck = C.make_checkpoint(sd_c, {"module." + k: v for k, v in sd_f.items()}, global_step=1234)   # DataParallel prefix
    buf = io.BytesIO()
    torch.save(ck, buf)
    buf.seek(0)
    got = C.read_checkpoint(torch.load(buf, map_location="cpu", weights_only=True))
    assert got["global_step"] == 1234 and got["cond"] is False and got["optimizer"] is None
    for k in sd_c:
        assert torch.equal(got["coarse"][k], sd_c[k]) and torch.equal(got["fine"][k], sd_f[k])


def test_bad_checkpoints_are_rejected():
    sd = init_state_dict(0)
    with pytest.raises(ValueError):
        C.read_checkpoint({"foo": 1})
    bad = dict(sd)
    bad.pop("rgb_linear.bias")
    with pytest.raises(ValueError):
        C.validate_state_dict(bad)
    bad = dict(sd)
    bad["pts_linears.5.weight"] = torch.zeros(256, 300)
    with pytest.raises(ValueError):
        C.validate_state_dict(bad)
    with pytest.raises(ValueError):
        C.make_checkpoint(sd, init_state_dict(1, cond=True)) and C.read_checkpoint(C.make_checkpoint(sd, init_state_dict(1, cond=True)))


class _FakeTrainer:
    def __init__(self, n):
        from fashion_nerf_b200.train import FlatAdam
        self.opt = FlatAdam(n, "cpu")


def test_canonical_torch_adam_state_is_converted_to_the_flat_layout():
    """A checkpoint written by the canonical training loop carries torch.optim.Adam's {'state','param_groups'} over
    list(model.parameters()) + list(model_fine.parameters()); restore_optimizer maps it onto the flat buffers
    (module order -> flat order differs: views_linears comes before feature/alpha in the module)."""
    torch.manual_seed(1)
    mc, mf = CanonicalNeRF(), CanonicalNeRF()
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=3e-4, betas=(0.8, 0.95), eps=1e-7)
    for _ in range(3):
        for p in list(mc.parameters()) + list(mf.parameters()):
            p.grad = torch.randn_like(p)
        opt.step()
    ck = {"global_step": 3, "network_fn_state_dict": mc.state_dict(), "network_fine_state_dict": mf.state_dict(),
          "optimizer_state_dict": opt.state_dict()}
    got = C.read_checkpoint(ck)
    n = sum(p.numel() for p in mc.parameters())
    tr = _FakeTrainer(2 * n)
    C.restore_optimizer(tr, got)
    assert tr.opt.t == 3 and tr.opt.lr == 3e-4 and (tr.opt.b1, tr.opt.b2) == (0.8, 0.95) and tr.opt.eps == 1e-7
    # the flat layout is pts 0..7, alpha, feature, views, rgb: compare tensor by tensor through unflatten
    for net, (module, lo) in enumerate(((mc, 0), (mf, n))):
        moments = unflatten(tr.opt.m[lo:lo + n])
        sq = unflatten(tr.opt.v[lo:lo + n])
        for name, p in module.named_parameters():
            st = opt.state[p]
            assert torch.equal(moments[name], st["exp_avg"]) and torch.equal(sq[name], st["exp_avg_sq"]), (net, name)
    # a malformed optimizer dictionary is a clear ValueError, not a KeyError
    with pytest.raises(ValueError):
        C.restore_optimizer(tr, {"optimizer": {"foo": 1}, "cond": False, "fine": None})
    with pytest.raises(ValueError):
        C.restore_optimizer(_FakeTrainer(n), got)          # one network's worth of state buffers
