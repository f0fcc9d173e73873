"""Launch the bf16 network query a few times with the library given in FNERF_LIB (for ncu A/B runs)."""
import os, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
if os.environ.get("FNERF_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["FNERF_LIB"])
F.load_library()
dev = torch.device('cuda:0')
net = F.NerfNetwork.random(1, dev)
R, S = 16384, 192
g = torch.Generator().manual_seed(1)
o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
vd, dn = F.ops.ray_setup(d)
for _ in range(6):
    raw = F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
torch.cuda.synchronize()
print("ok")
