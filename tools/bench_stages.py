"""Standalone HBM-roofline benches of the sampling / compositing kernels (R = 2^20 rays,
raw ~ N(0,1), working set >> L2).  Prints one JSON object; algorithmic bytes per ray from SURVEY.md 8(d)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F
dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
R = 1 << 20


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"R": R, "hbm_peak_gbs": HBM, "peak_source": "measured" if "when" in peaks else "fallback"}
g = torch.Generator(device="cuda").manual_seed(0)
for S in (64, 192, 1024):
    z = torch.sort(torch.rand(R, S, device=dev, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, device=dev, generator=g)
    dn = torch.ones(R, device=dev)
    ms = timeit(lambda: F.ops.composite_fwd(raw, z, dn))
    b = (24 * S + 36) * R
    out[f"composite_fwd_S{S}"] = {"ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / HBM, "bytes_per_ray": 24 * S + 36}
    g_rgb = torch.randn(R, 3, device=dev, generator=g)
    ms = timeit(lambda: F.ops.composite_bwd(raw, z, dn, g_rgb))
    b = (36 * S + 24) * R
    out[f"composite_bwd_S{S}"] = {"ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / HBM, "bytes_per_ray": 36 * S + 24}
    del z, raw
near, far = torch.full((R,), 2.0, device=dev), torch.full((R,), 6.0, device=dev)
t = torch.linspace(0, 1, 64).to(dev)
u = torch.rand(R, 64, device=dev, generator=g)
ms = timeit(lambda: F.ops.stratified(near, far, t, u))
out["stratified_N64"] = {"ms": ms, "GBps": 520 * R / ms / 1e6, "frac": 520 * R / ms / 1e6 / HBM, "bytes_per_ray": 520}
zc = F.ops.stratified(near, far, t, u)
w = torch.rand(R, 64, device=dev, generator=g)
uf = torch.rand(R, 128, device=dev, generator=g)
ms = timeit(lambda: F.ops.importance(zc, w, uf, want_idx=False))
out["importance_64_128"] = {"ms": ms, "GBps": 1796 * R / ms / 1e6, "frac": 1796 * R / ms / 1e6 / HBM, "bytes_per_ray": 1796}
print(json.dumps(out, indent=1))
