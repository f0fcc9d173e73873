"""Randomised soak of the register-path importance kernels and the compositing kernels against the oracle: many random
launch sizes (odd R, more rays than warps in the grid), every dispatched shape, sorted / unsorted / shared uniform rows,
peaky weights.  Bit-exact for importance, <= 1e-5 for compositing.  python tests/debug/soak_sampling.py [iterations]"""
import os, sys, random, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import fashion_nerf_b200 as F
from oracle import nerf_oracle as O
dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rnd = random.Random(1234)
shapes = [(64, 128), (32, 32), (64, 64), (128, 128), (128, 256), (256, 768)]
bad = 0
for it in range(iters):
    Nc, Nf = shapes[it % len(shapes)]
    R = rnd.choice([1, 2, 3, 17, 255, 777, 4096, 9473, 18945, 30001]) if Nf < 768 else rnd.choice([1, 5, 301, 4737, 9475])
    g = torch.Generator().manual_seed(it)
    near, far = torch.full((R,), 2.0), torch.full((R,), 6.0)
    z = O.stratified(near, far, torch.linspace(0, 1, Nc), torch.rand(R, Nc, generator=g))
    raw = torch.randn(R, Nc, 4, generator=g) * rnd.choice([1.0, 4.0, 12.0])
    w = O.raw2outputs(raw, z, torch.ones(R))["weights"]
    kind = rnd.choice(["random", "row", "mixed", "dups"])
    if kind == "row":
        u_dev = torch.linspace(0, 1, Nf); u = u_dev[None].expand(R, Nf)
    else:
        u = torch.rand(R, Nf, generator=g)
        if kind == "mixed": u[::2] = torch.sort(u[::2], -1)[0]
        if kind == "dups": u[:, 1::2] = u[:, 0::2]
        u_dev = u
    ref = O.sample_pdf(z, w, u)
    got = F.ops.importance(z.to(dev), w.to(dev), u_dev.contiguous().to(dev))
    ok = torch.equal(got["inds"].cpu().long(), ref["inds"]) and torch.equal(got["z_samples"].cpu(), ref["z_samples"]) and torch.equal(got["z_f"].cpu(), ref["z_f"])
    bad += not ok
    print(f"importance {Nc:3d}+{Nf:3d} R={R:6d} {kind:6s} {'ok' if ok else 'MISMATCH'}", flush=True)
for it in range(iters // 2):
    S = rnd.choice([1, 31, 32, 33, 64, 100, 192, 256, 257, 300, 512, 777, 1024, 1100])
    R = rnd.choice([1, 2, 33, 1000, 4737, 9473, 20011, 40003]) if S <= 256 else rnd.choice([1, 3, 100, 4737, 6001])
    white = rnd.random() < 0.5
    g = torch.Generator().manual_seed(1000 + it)
    zz = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
    raw = torch.randn(R, S, 4, generator=g)
    dn = 1 + torch.rand(R, generator=g)
    ref = O.raw2outputs(raw, zz, dn, white)
    got = F.ops.composite_fwd(raw.to(dev), zz.to(dev), dn.to(dev), white_bkgd=white)
    e = max((got[k].cpu() - ref[k]).abs().max().item() for k in ("rgb", "acc", "weights"))
    g_rgb = torch.randn(R, 3, generator=g)
    want = O.composite_bwd(raw.double(), zz.double(), dn.double(), g_rgb.double(), torch.zeros(R).double(), torch.zeros(R).double(), white)
    ref32 = O.composite_bwd(raw, zz, dn, g_rgb, torch.zeros(R), torch.zeros(R), white)
    gb = F.ops.composite_bwd(raw.to(dev), zz.to(dev), dn.to(dev), g_rgb.to(dev), white_bkgd=white).cpu()
    eb = (gb.double() - want).abs().max().item()
    tol = max(4 * (ref32.double() - want).abs().max().item(), 1e-5 * max(want.abs().max().item(), 1.0))
    ok = e <= 1e-5 and eb <= tol and bool(torch.isfinite(gb).all())
    bad += not ok
    print(f"composite S={S:4d} R={R:6d} white={int(white)} fwd err {e:.2e} bwd err {eb:.2e} (tol {tol:.2e}) {'ok' if ok else 'MISMATCH'}", flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
