"""CPU-side checks of the drop-in boundary: libfnerf.so loads without a GPU, exports every symbol
include/fnerf.h declares, validates arguments before any launch, and the Python layer refuses CPU
tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import fashion_nerf_b200 as F
    if not os.path.exists(F._lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return F.load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "fnerf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fnerf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from fashion_nerf_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fnerf.h but not exported"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_sizes(lib):
    assert lib.fnerf_abi_version() == 4
    assert lib.fnerf_param_count(0) == 595_844
    assert lib.fnerf_param_count(1) == 595_844 + 256 * 256         # A.8
    assert lib.fnerf_packed_bytes(0) % 256 == 0 and lib.fnerf_packed_bytes(0) > 1_190_000
    assert lib.fnerf_render_rays_workspace_bytes(4096, 64, 128) > 4096 * 192 * 16
    assert lib.fnerf_render_rays_workspace_bytes(-1, 64, 128) == -1


def test_argument_validation_happens_before_any_launch(lib):
    """Negative codes come from host-side validation, so they are observable without a GPU."""
    assert lib.fnerf_stratified(None, None, None, None, None, 4, 4, 0, None) == -1
    assert b"null" in lib.fnerf_last_error()
    assert lib.fnerf_stratified(None, None, None, None, None, 4, 0, 0, None) == -2           # N < 1
    assert lib.fnerf_importance(None, None, None, 0, None, None, None, None, 4, 2, 4, None) == -2   # Nc < 3
    assert lib.fnerf_importance(None, None, None, 3, None, None, None, None, 4, 8, 4, None) == -4   # bad stride
    assert lib.fnerf_composite_fwd(None, None, None, None, None, None, None, None, None, 1, 0, 0, None) == -2
    assert lib.fnerf_mlp_fwd(7, None, 0, None, None, None, None, None, None, 0, None, 1, 1, None) == -4
    assert lib.fnerf_render_rays(None, None) == -1
    # empty inputs are a no-op success
    assert lib.fnerf_stratified(None, None, None, None, None, 0, 4, 0, None) == 0
    assert lib.fnerf_composite_fwd(None, None, None, None, None, None, None, None, None, 0, 8, 0, None) == 0
    # misaligned raw pointer is rejected
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf) + 4
    assert lib.fnerf_composite_fwd(p, p, p, None, p, p, p, p, None, 1, 1, 0, None) == -3


def test_render_args_struct_matches_header():
    from fashion_nerf_b200._lib import RenderArgs
    text = open(os.path.join(ROOT, "include", "fnerf.h")).read()
    body = text[text.index("typedef struct fnerf_render_args {"):text.index("} fnerf_render_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        for piece in stmt.split(","):          # "int64_t R, Nc, Nf" declares three fields
            if piece.strip():
                fields.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", piece.strip())[0])
    assert fields == [f[0] for f in RenderArgs._fields_]


def test_no_cpu_fallback():
    import fashion_nerf_b200 as F
    with pytest.raises(F.FnerfError):
        F.ops.ray_setup(torch.zeros(4, 3))
    with pytest.raises(F.FnerfError):
        F.render_rays(None, torch.zeros(4, 3), torch.zeros(4, 3), 2.0, 6.0, 8, 8)
    with pytest.raises(RuntimeError):
        F.NerfNetwork(torch.zeros(595_844))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fashion_nerf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle's bits", ""), fn


def test_flat_layout_round_trip_on_cpu():
    import fashion_nerf_b200 as F
    for cond in (False, True):
        sd = F.init_state_dict(5, cond)
        flat = F.flatten_state_dict(sd, cond)
        back = F.unflatten(flat, cond)
        assert all(torch.equal(back[k], sd[k]) for k in sd)
