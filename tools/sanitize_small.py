"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): inference render through the fused query + compositing
kernel (64 + 128 samples, tile groups of 1 and 3, ragged tail), then training render + backward (tape forward with its
tape-writer warps, pipelined backward), bf16 and fp32."""
import sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
model = F.NerfModel.random(dev)
o, d = F.pinhole_rays(9, 9)
o, d = o.to(dev), d.to(dev)
with torch.no_grad():
    out = F.render_rays(model, o, d, 2.0, 6.0, 64, 128)
torch.cuda.synchronize()
print("fused inference ok", float(out["rgb"].abs().sum()))
for prec in ("bf16", "fp32"):
    model.coarse.flat.requires_grad_(True); model.fine.flat.requires_grad_(True)
    out = F.render_rays(model, o, d, 2.0, 6.0, 16, 16, precision=prec)
    (out["rgb"].sum() + out["rgb0"].sum()).backward()
    torch.cuda.synchronize()
    print(prec, "ok", float(out["rgb"].abs().sum()))
