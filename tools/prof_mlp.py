"""Small driver for ncu: a few launches of the bf16 tcgen05 network query (16384 rays x 192 samples)
and of the compositing / sampling kernels at the standalone-roofline sizes."""
import sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
net = F.NerfNetwork.random(1, dev)
R, S = 16384, 192
g = torch.Generator().manual_seed(1)
o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
vd, dn = F.ops.ray_setup(d)
for _ in range(4):
    raw = F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
torch.cuda.synchronize()
Rb = 1 << 18
zb = torch.sort(torch.rand(Rb, S, device=dev) * 4 + 2, -1)[0]
rawb = torch.randn(Rb, S, 4, device=dev)
dnb = torch.ones(Rb, device=dev)
for _ in range(3):
    out = F.ops.composite_fwd(rawb, zb, dnb)
torch.cuda.synchronize()
print("ok")
