"""Checkpoint format: the canonical nerf-pytorch dictionary, so externally trained NeRFs render
through this path and checkpoints written here load into the canonical `NeRF(D=8, W=256, skips=[4],
use_viewdirs=True)` module unchanged.

    {"global_step": int,
     "network_fn_state_dict":   {pts_linears.{0..7}.weight/bias, views_linears.0.*, feature_linear.*, alpha_linear.*, rgb_linear.*},
     "network_fine_state_dict": same keys (absent when coarse and fine share one network),
     "optimizer_state_dict":    {"step", "exp_avg", "exp_avg_sq", "lr", "betas", "eps"} over the flat parameter layout
                                (this package's Adam).  A canonical torch.optim.Adam state dict ({"state", "param_groups"}
                                over list(model.parameters()) + list(model_fine.parameters())) is recognised on load and
                                converted; checkpoints WRITTEN here carry the flat form, so only the network
                                dictionaries are loadable by the canonical code unchanged.}

Conditioned networks (A.8) differ only in `pts_linears.5.weight` being [256, 63+256+256]; `cond` is inferred
from that shape on load.  Everything here is host-side dictionary work and runs without a GPU; only
`load_model` builds device objects.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .model import NerfModel, NerfNetwork, flatten_state_dict, layer_shapes, unflatten

_PREFIXES = ("module.", "_orig_mod.")     # DataParallel / torch.compile wrappers of the canonical module


def param_shapes(cond: bool = False) -> Dict[str, tuple]:
    """Tensor shapes of the canonical module's state dict, in flat order."""
    out = {}
    for name, (o, i) in layer_shapes(cond).items():
        out[name + ".weight"] = (o, i)
        out[name + ".bias"] = (o,)
    return out


def _clean(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for p in _PREFIXES:
            if k.startswith(p):
                k = k[len(p):]
        out[k] = v.detach().to("cpu", torch.float32)
    return out


def infer_cond(sd: Dict[str, torch.Tensor]) -> bool:
    """True when the state dict is the conditioned variant (layer-5 fan-in 575 instead of 319)."""
    w5 = _clean(sd)["pts_linears.5.weight"]
    if tuple(w5.shape) == param_shapes(False)["pts_linears.5.weight"]:
        return False
    if tuple(w5.shape) == param_shapes(True)["pts_linears.5.weight"]:
        return True
    raise ValueError(f"pts_linears.5.weight has shape {tuple(w5.shape)}: not an 8x256 skip-4 NeRF")


def validate_state_dict(sd: Dict[str, torch.Tensor], cond: Optional[bool] = None) -> Tuple[Dict[str, torch.Tensor], bool]:
    """Checks names and shapes against the architecture; returns (cleaned fp32 CPU dict, cond)."""
    sd = _clean(sd)
    if cond is None:
        cond = infer_cond(sd)
    shapes = param_shapes(cond)
    missing = [k for k in shapes if k not in sd]
    extra = [k for k in sd if k not in shapes]
    if missing or extra:
        raise ValueError(f"state dict mismatch: missing {missing}, unexpected {extra}")
    for k, shp in shapes.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(shp)}")
    return {k: sd[k] for k in shapes}, cond


def make_checkpoint(coarse_sd, fine_sd=None, *, global_step: int = 0, optimizer=None) -> dict:
    """Builds the checkpoint dictionary from state dicts (and optionally a train.FlatAdam)."""
    ck = {"global_step": int(global_step), "network_fn_state_dict": validate_state_dict(coarse_sd)[0]}
    if fine_sd is not None:
        ck["network_fine_state_dict"] = validate_state_dict(fine_sd)[0]
    if optimizer is not None:
        ck["optimizer_state_dict"] = {"step": int(optimizer.t), "exp_avg": optimizer.m.detach().cpu().clone(),
                                      "exp_avg_sq": optimizer.v.detach().cpu().clone(),
                                      "lr": optimizer.lr, "betas": (optimizer.b1, optimizer.b2), "eps": optimizer.eps}
    return ck


def save_checkpoint(path: str, model: NerfModel, *, global_step: int = 0, trainer=None) -> None:
    fine = None if model.fine is model.coarse else model.fine.state_dict()
    torch.save(make_checkpoint(model.coarse.state_dict(), fine, global_step=global_step,
                               optimizer=None if trainer is None else trainer.opt), path)


def read_checkpoint(path_or_dict) -> dict:
    """Loads and validates a checkpoint; returns {"coarse", "fine" (or None), "cond", "global_step", "optimizer"}."""
    ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=True)
    if "network_fn_state_dict" not in ck:
        raise ValueError("not a NeRF checkpoint: no 'network_fn_state_dict'")
    coarse, cond = validate_state_dict(ck["network_fn_state_dict"])
    fine = None
    if ck.get("network_fine_state_dict") is not None:
        fine, cond_f = validate_state_dict(ck["network_fine_state_dict"])
        if cond_f != cond:
            raise ValueError("coarse and fine networks disagree on conditioning")
    return {"coarse": coarse, "fine": fine, "cond": cond, "global_step": int(ck.get("global_step", 0)),
            "optimizer": ck.get("optimizer_state_dict")}


def load_model(path_or_dict, device) -> Tuple[NerfModel, dict]:
    """Checkpoint -> NerfModel on `device` (packs the kernel blobs) + the parsed checkpoint."""
    ck = read_checkpoint(path_or_dict)
    coarse = NerfNetwork.from_state_dict(ck["coarse"], device, ck["cond"])
    fine = None if ck["fine"] is None else NerfNetwork.from_state_dict(ck["fine"], device, ck["cond"])
    return NerfModel(coarse, fine), ck


# parameter order of the canonical module (its __init__ declares pts_linears, views_linears, feature_linear,
# alpha_linear, rgb_linear): torch.optim.Adam numbers its state by position in list(model.parameters())
_MODULE_ORDER = [f"pts_linears.{i}" for i in range(8)] + ["views_linears.0", "feature_linear", "alpha_linear", "rgb_linear"]


def convert_torch_adam_state(opt_sd: dict, cond: bool, two_networks: bool) -> dict:
    """torch.optim.Adam.state_dict() of Adam(list(coarse.parameters()) + list(fine.parameters())) -> the flat form."""
    state, groups = opt_sd["state"], opt_sd["param_groups"]
    names = [n + sfx for n in _MODULE_ORDER for sfx in (".weight", ".bias")]
    shapes = param_shapes(cond)
    n_nets = 2 if two_networks else 1
    order = [pid for g in groups for pid in g["params"]]
    if len(order) != n_nets * len(names):
        raise ValueError(f"torch Adam state covers {len(order)} tensors, expected {n_nets * len(names)} "
                         f"({n_nets} network(s) x {len(names)} parameters)")
    nets, step = [], 0
    for k in range(n_nets):
        avg, sq = {}, {}
        for j, name in enumerate(names):
            st = state.get(order[k * len(names) + j])
            if st is None:                                   # parameter never stepped: zero moments
                avg[name], sq[name] = torch.zeros(shapes[name]), torch.zeros(shapes[name])
                continue
            if tuple(st["exp_avg"].shape) != tuple(shapes[name]):
                raise ValueError(f"torch Adam state: {name} has shape {tuple(st['exp_avg'].shape)} != {tuple(shapes[name])}")
            avg[name], sq[name] = st["exp_avg"].detach().float().cpu(), st["exp_avg_sq"].detach().float().cpu()
            step = max(step, int(st["step"]))
        nets.append((flatten_state_dict(avg, cond), flatten_state_dict(sq, cond)))
    g0 = groups[0]
    return {"step": step, "exp_avg": torch.cat([a for a, _ in nets]), "exp_avg_sq": torch.cat([b for _, b in nets]),
            "lr": float(g0["lr"]), "betas": tuple(g0["betas"]), "eps": float(g0["eps"])}


def restore_optimizer(trainer, ck: dict, *, restore_hyperparameters: bool = True) -> None:
    """Puts a checkpoint's Adam moments / step count (and lr / betas / eps when recorded) back into a train.Trainer.
    Accepts this package's flat form and a canonical torch.optim.Adam state dict."""
    st = ck.get("optimizer")
    if st is None:
        return
    if "state" in st and "param_groups" in st:
        st = convert_torch_adam_state(st, ck["cond"], ck["fine"] is not None)
    for k in ("step", "exp_avg", "exp_avg_sq"):
        if k not in st:
            raise ValueError(f"optimizer_state_dict has no '{k}': neither this package's flat Adam state nor a torch.optim.Adam one")
    if st["exp_avg"].numel() != trainer.opt.m.numel() or st["exp_avg_sq"].numel() != trainer.opt.v.numel():
        raise ValueError("optimizer state does not match the model's parameter count")
    trainer.opt.m.copy_(st["exp_avg"])
    trainer.opt.v.copy_(st["exp_avg_sq"])
    trainer.opt.t = int(st["step"])
    if restore_hyperparameters:
        if "lr" in st:
            trainer.opt.lr = float(st["lr"])
        if "betas" in st:
            trainer.opt.b1, trainer.opt.b2 = (float(b) for b in st["betas"])
        if "eps" in st:
            trainer.opt.eps = float(st["eps"])


__all__ = ["save_checkpoint", "read_checkpoint", "load_model", "make_checkpoint", "validate_state_dict", "infer_cond",
           "restore_optimizer", "convert_torch_adam_state", "flatten_state_dict", "unflatten"]
