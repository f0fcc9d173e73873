"""Times fnerf_importance at R = 2^20 rays for the two bench shapes, with per-ray random uniforms and with the shared
deterministic row (sort skipped).  Prints one JSON object; bytes per ray as in bench.py's roofline_stages."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F
dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
R = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, iters=5, warm=2):
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


out = {}
near, far = torch.full((R,), 2.0, device=dev), torch.full((R,), 6.0, device=dev)
for nc, nf in ((64, 128), (256, 768)):
    t = torch.linspace(0, 1, nc).to(dev)
    zc = F.ops.stratified(near, far, t, torch.rand(R, nc, device=dev, generator=g))
    w = torch.rand(R, nc, device=dev, generator=g)
    u = torch.rand(R, nf, device=dev, generator=g)
    b = 8 * nc + 4 * nf + 4 * (nc + nf) + 4
    for name, uu in (("random", u), ("linspace_row", torch.linspace(0, 1, nf).to(dev))):
        ms = timeit(lambda: F.ops.importance(zc, w, uu, want_idx=False))
        out[f"importance_{nc}_{nf}_{name}"] = {"ms": round(ms, 4), "GBps": round(b * R / ms / 1e6, 1),
                                               "frac": round(b * R / ms / 1e6 / peaks["hbm_gbs"], 4), "bytes_per_ray": b}
    del zc, w, u
print(json.dumps(out, indent=1))
