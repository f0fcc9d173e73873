import torch, sys, os
sys.path.insert(0, '/root/repo')
import fashion_nerf_b200 as F
dev = torch.device("cuda:0")
model = F.NerfModel.random(dev)
o_all, d_all = F.pinhole_rays(800, 800)
n = int(os.environ.get("FP32_RAYS", "80000"))
o, d = o_all[:n].to(dev), d_all[:n].to(dev)
g = torch.Generator(device=dev).manual_seed(0)
us, uf = torch.rand(n, 64, device=dev, generator=g), torch.rand(n, 128, device=dev, generator=g)
def step():
    with torch.no_grad():
        F.render_rays(model, o, d, 2.0, 6.0, 64, 128, u_strat=us, u_fine=uf, precision="fp32")
step(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"fp32 path: {ms:.1f} ms, {n/ms/1e3:.4f} Mrays/s, {n*1186816*256/ms/1e9:.2f} TFLOP/s")
