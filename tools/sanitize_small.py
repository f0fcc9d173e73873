"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): render + backward on 64 rays."""
import sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
model = F.NerfModel.random(dev)
o, d = F.pinhole_rays(8, 8)
o, d = o.to(dev), d.to(dev)
for prec in ("bf16", "fp32"):
    model.coarse.flat.requires_grad_(True); model.fine.flat.requires_grad_(True)
    out = F.render_rays(model, o, d, 2.0, 6.0, 16, 16, precision=prec)
    (out["rgb"].sum() + out["rgb0"].sum()).backward()
    torch.cuda.synchronize()
    print(prec, "ok", float(out["rgb"].abs().sum()))
