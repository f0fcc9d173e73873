"""fashion_nerf_b200: B200-native NeRF render/train hot path (sm_100a CUDA behind a C ABI).

Public surface:
  render_rays(model, rays_o, rays_d, near, far, N_samples, N_importance, cond=None, ...)
  render_image(...)            full-frame chunked driver
  NerfModel / NerfNetwork      fp32 master parameters + packed kernel blobs
  checkpoint.*                 nerf-pytorch style checkpoint dictionaries <-> NerfModel
  ops.*                        stage-level operators (stratified, importance, posenc, mlp_fwd, composite_*)
Importing the package does not load CUDA; the first operator call loads libfnerf.so and fails
loudly if it is missing (no CPU fallback).
"""
from . import checkpoint, ops  # noqa: F401
from ._lib import FnerfError, load as load_library  # noqa: F401
from .model import NerfModel, NerfNetwork, flatten_state_dict, init_state_dict, unflatten  # noqa: F401
from .rays import pinhole_rays  # noqa: F401
from .render import render_image, render_rays  # noqa: F401

__all__ = ["render_rays", "render_image", "NerfModel", "NerfNetwork", "ops", "checkpoint", "pinhole_rays", "load_library",
           "FnerfError", "flatten_state_dict", "init_state_dict", "unflatten"]
