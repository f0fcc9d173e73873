"""One launch of every memory-bound stage kernel at R = 2^20 rays (SURVEY.md 8d sizes), for an `ncu --set full` capture:
   ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_importance|k_stratified' -o gpurun_out/stages python tools/stage_ncu.py
Launch order: stratified 64, importance 64/128, importance 256/768, then per S in (64, 192, 1024): composite fwd, composite bwd."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F
dev = torch.device("cuda:0")
R = 1 << 20
g = torch.Generator(device="cuda").manual_seed(0)
near, far = torch.full((R,), 2.0, device=dev), torch.full((R,), 6.0, device=dev)
dn = 1.0 + torch.rand(R, device=dev, generator=g)
t64 = torch.linspace(0, 1, 64).to(dev)
u = torch.rand(R, 64, device=dev, generator=g)
zc = F.ops.stratified(near, far, t64, u)
w = torch.rand(R, 64, device=dev, generator=g)
uf = torch.rand(R, 128, device=dev, generator=g)
F.ops.importance(zc, w, uf, want_idx=False)
del u, zc, w, uf
t256 = torch.linspace(0, 1, 256).to(dev)
zc = torch.cumsum(torch.rand(R, 256, device=dev, generator=g), -1) * (4.0 / 256) + 2.0
w = torch.rand(R, 256, device=dev, generator=g)
uf = torch.rand(R, 768, device=dev, generator=g)
F.ops.importance(zc, w, uf, want_idx=False)
del zc, w, uf
for S in (64, 192, 1024):
    raw = torch.randn(R, S, 4, device=dev, generator=g)
    z = torch.cumsum(torch.rand(R, S, device=dev, generator=g), -1) * (4.0 / S) + 2.0
    g_rgb = torch.randn(R, 3, device=dev, generator=g)
    F.ops.composite_fwd(raw, z, dn)
    F.ops.composite_bwd(raw, z, dn, g_rgb)
    del raw, z, g_rgb
torch.cuda.synchronize()
print("ok")
