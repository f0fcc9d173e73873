"""ctypes binding of libfnerf.so (include/fnerf.h).  There is no fallback: if the shared library is
missing or a call fails, this raises -- the product path never routes through a CPU/PyTorch
re-implementation."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FNERF_LIB") or os.path.join(HERE, "libfnerf.so")   # FNERF_LIB: A/B builds (tools/)

ABI_VERSION = 4
PRECISION_FP32 = 0
PRECISION_BF16 = 1


class FnerfError(RuntimeError):
    pass


class RenderArgs(ctypes.Structure):
    """Mirror of ``fnerf_render_args`` (include/fnerf.h)."""
    _fields_ = [
        ("packed_coarse", c_void_p), ("packed_fine", c_void_p),
        ("cond", c_int), ("precision", c_int),
        ("rays_o", c_void_p), ("rays_d", c_void_p), ("near", c_void_p), ("far", c_void_p),
        ("t_vals", c_void_p), ("u_strat", c_void_p), ("u_fine", c_void_p),
        ("u_fine_row_stride", c_int64),
        ("raw_noise_coarse", c_void_p), ("raw_noise_fine", c_void_p),
        ("cond_proj_coarse", c_void_p), ("cond_proj_fine", c_void_p), ("cond_index", c_void_p),
        ("C", c_int64), ("R", c_int64), ("Nc", c_int64), ("Nf", c_int64),
        ("white_bkgd", c_int), ("lindisp", c_int),
        ("rgb", c_void_p), ("disp", c_void_p), ("acc", c_void_p), ("depth", c_void_p),
        ("rgb0", c_void_p), ("disp0", c_void_p), ("acc0", c_void_p), ("z_std", c_void_p),
        ("depth0", c_void_p),
        ("z_c", c_void_p), ("z_f", c_void_p), ("raw_c", c_void_p), ("raw_f", c_void_p),
        ("weights_c", c_void_p), ("weights_f", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("ev_coarse_start", c_void_p), ("ev_coarse_stop", c_void_p),
        ("ev_fine_start", c_void_p), ("ev_fine_stop", c_void_p),
        ("tape_coarse", c_void_p), ("tape_coarse_bytes", c_int64),
        ("tape_fine", c_void_p), ("tape_fine_bytes", c_int64),
        ("fuse_composite", c_int),
    ]


# name -> (restype, argtypes); every symbol include/fnerf.h declares
SIGNATURES = {
    "fnerf_abi_version": (c_int, []),
    "fnerf_last_error": (c_char_p, []),
    "fnerf_launch_count": (c_int64, []),
    "fnerf_param_count": (c_int64, [c_int]),
    "fnerf_packed_bytes": (c_int64, [c_int]),
    "fnerf_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "fnerf_unpack_weights": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "fnerf_ray_setup": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fnerf_stratified": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "fnerf_importance": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int64, c_int64, c_void_p]),
    "fnerf_posenc": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "fnerf_cond_project": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fnerf_mlp_fwd": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64, c_void_p, c_int64, c_int64, c_void_p]),
    "fnerf_mlp_fwd_composite_supported": (c_int, [c_int64]),
    "fnerf_mlp_fwd_composite": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int, c_void_p]),
    "fnerf_mlp_bwd_workspace_bytes": (c_int64, [c_int, c_int, c_int64, c_int64]),
    "fnerf_mlp_bwd": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "fnerf_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.c_float, c_int64, ctypes.c_float, c_void_p]),
    "fnerf_allreduce_adam_step": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_float, c_int64, ctypes.c_float, c_void_p]),
    "fnerf_multimem_allreduce": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p]),
    "fnerf_mlp_tape_bytes": (c_int64, [c_int64, c_int64]),
    "fnerf_mlp_fwd_tape": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "fnerf_mlp_bwd_tape_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "fnerf_mlp_bwd_tape": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                   c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "fnerf_composite_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "fnerf_composite_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int, c_void_p]),
    "fnerf_render_rays_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "fnerf_render_rays": (c_int, [ctypes.POINTER(RenderArgs), c_void_p]),
    "fnerf_debug_pipe_stats": (c_int, [c_void_p]),
    "fnerf_debug_fdiv_mismatches": (c_int, [c_int64, ctypes.c_uint64, c_void_p, c_void_p]),
    "fnerf_debug_wgrad_tc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int64, c_int, c_int64, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libfnerf.so (built in-tree by ``fashion_nerf_b200._build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FnerfError(f"{LIB_PATH} is missing: run `python -m fashion_nerf_b200._build` "
                         "(or __graft_entry__.build()); there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.fnerf_abi_version()
    if v != ABI_VERSION:
        raise FnerfError(f"libfnerf.so ABI {v} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fnerf_last_error()
        raise FnerfError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
