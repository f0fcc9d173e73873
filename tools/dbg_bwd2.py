"""Debug: bf16 backward vs fp32 SGEMM backward fed with the SAME bf16-rounded weights (mask-flip free check is
not possible from outside, so this prints the error profile per tensor and sample-count dependence)."""
import sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
for R, S in ((40, 24), (400, 64), (4000, 64)):
    g = torch.Generator().manual_seed(21)
    o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
    g_raw = torch.randn(R, S, 4, generator=g).to(dev)
    net = F.NerfNetwork.random(4, dev)
    # weights rounded to bf16 so both paths hold identical weights (biases stay fp32)
    sd = F.unflatten(net.flat, False)
    sd = {k: (v.to(torch.bfloat16).float() if k.endswith("weight") else v) for k, v in sd.items()}
    net = F.NerfNetwork.from_state_dict(sd, dev, cond=False)
    vd, _ = F.ops.ray_setup(d)
    ref = torch.zeros(net.flat.numel(), device=dev); got = torch.zeros_like(ref)
    F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, ref, precision="fp32")
    F.ops.mlp_bwd(net.packed, o, d, vd, z, g_raw, got, precision="bf16")
    sr, sg = F.unflatten(ref, False), F.unflatten(got, False)
    print(R, S, " ".join(f"{((sg[k]-sr[k]).norm()/sr[k].norm()).item():.3f}" for k in sr))
