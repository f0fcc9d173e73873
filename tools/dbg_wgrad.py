"""Debug: tensor-core wgrad (MN-major operands from tape images) vs torch matmul."""
import ctypes, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
lib = F.load_library()
dev = torch.device('cuda:0')


def to_images(t):
    """[T*128, 64*nkb] float -> uint8 image buffer [T, nkb, 128 rows, 128 B] with the 128-byte swizzle."""
    M, C = t.shape
    T, nkb = M // 128, C // 64
    b = t.to(torch.bfloat16).reshape(T, 128, nkb, 8, 8).permute(0, 2, 1, 3, 4).contiguous()   # [T, kb, r, chunk, 8]
    r = torch.arange(128, device=t.device)
    c = torch.arange(8, device=t.device)
    src = (c[None, :] ^ (r[:, None] & 7))                      # out[r, p] = in[r, p ^ (r&7)]
    idx = src[None, None, :, :, None].expand(T, nkb, 128, 8, 8)
    return torch.gather(b, 3, idx).contiguous().view(torch.uint8).reshape(-1)


for T, n_kb, x_kb in ((1, 4, 4), (3, 4, 4), (5, 2, 1), (300, 4, 4), (7, 4, 1)):
    g = torch.Generator(device="cpu").manual_seed(T)
    dz = (torch.randn(T * 128, 64 * n_kb, generator=g) * 0.5).to(dev)
    x = (torch.randn(T * 128, 64 * x_kb, generator=g) * 0.5).to(dev)
    dzi, xi = to_images(dz), to_images(x)
    dw = torch.zeros(64 * n_kb, 64 * x_kb, device=dev)
    rc = lib.fnerf_debug_wgrad_tc(ctypes.c_void_p(dzi.data_ptr()), n_kb, ctypes.c_void_p(xi.data_ptr()), x_kb,
                                  ctypes.c_void_p(dw.data_ptr()), ctypes.c_int64(64 * x_kb), 64 * x_kb, ctypes.c_int64(T),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = dz.to(torch.bfloat16).float().t() @ x.to(torch.bfloat16).float()
    err = (dw - ref).abs().max().item()
    print(f"T={T} n_kb={n_kb} x_kb={x_kb} rc={rc} max abs err {err:.3e} (ref max {ref.abs().max().item():.2f})", flush=True)
