import sys, torch
sys.path.insert(0, '.')
from oracle import nerf_oracle as O
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
_, d = O.pinhole_rays(37, 53)
d = torch.cat([d, torch.randn(1000, 3, generator=torch.Generator().manual_seed(0)) * 3])
vd_ref, dn_ref = O.ray_setup_exact(d)
vd, dn = F.ops.ray_setup(d.to(dev))
vd, dn = vd.cpu(), dn.cpu()
print("dn mismatches", (dn != dn_ref).sum().item(), "of", dn.numel(), "max abs", (dn - dn_ref).abs().max().item())
print("vd mismatches", (vd != vd_ref).sum().item(), "of", vd.numel(), "max abs", (vd - vd_ref).abs().max().item())
# alternatives on CPU
n2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
print("sqrt variants equal:", torch.equal(torch.sqrt(n2), dn_ref))
vd_alt = d * (1.0 / dn_ref[:, None])
print("vd == d*(1/dn) count mismatches vs gpu", (vd_alt != vd).sum().item())
vd_alt2 = d / dn[:, None]
print("cpu div using gpu dn mismatches vs gpu", (vd_alt2 != vd).sum().item())
print("---- deeper")
dg = d.to(dev)
n2g = ((dg[:, 0] * dg[:, 0] + dg[:, 1] * dg[:, 1]) + dg[:, 2] * dg[:, 2]).cpu()
print("n2 torch-gpu vs cpu mismatches", (n2g != n2).sum().item())
exact = torch.sqrt(n2.double()).float()
print("cpu sqrt vs correctly-rounded:", (torch.sqrt(n2) != exact).sum().item())
print("gpu kernel dn vs correctly-rounded:", (dn != exact).sum().item())
print("torch gpu sqrt vs correctly-rounded:", (torch.sqrt(n2.to(dev)).cpu() != exact).sum().item())
bad = (dn != dn_ref).nonzero().flatten()
for i in bad[:6].tolist():
    print(i, d[i].tolist(), n2[i].item(), dn[i].item(), dn_ref[i].item(), exact[i].item())
print(torch.__config__.show()[:600])
