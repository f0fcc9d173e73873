"""Debug: bf16 tensor-core MLP backward vs the fp32 SGEMM chain on the same inputs."""
import sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
R, S = int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 24
g = torch.Generator().manual_seed(21)
o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
g_raw = torch.randn(R, S, 4, generator=g).to(dev)
net = F.NerfNetwork.random(4, dev)
vd, _ = F.ops.ray_setup(d)
ref = torch.zeros(net.flat.numel(), device=dev); got = torch.zeros_like(ref)
F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, ref, precision="fp32")
torch.cuda.synchronize(); print("fp32 done", flush=True)
F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, got, precision="bf16")
torch.cuda.synchronize(); print("bf16 done", flush=True)
sr, sg = F.unflatten(ref, False), F.unflatten(got, False)
for k in sr:
    e = ((sg[k] - sr[k]).norm() / sr[k].norm().clamp_min(1e-12)).item()
    print(f"{k:28s} rel err {e:.3e}  |ref| {sr[k].norm().item():.3e} finite {bool(torch.isfinite(sg[k]).all())}")
if R * S >= 100000:
    for prec in ("bf16", "fp32"):
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, got, precision=prec); e1.record(); torch.cuda.synchronize()
        print(prec, "bwd ms", e0.elapsed_time(e1))
