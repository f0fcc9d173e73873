"""Render-kernel timing at the weight-sharing cluster size given by FNERF_MLP_CLUSTER (one process per setting);
prints TFLOP/s and a checksum of raw so settings can be compared bit-for-bit."""
import os, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
Rb, Sb = 16384, 192
gb = torch.Generator().manual_seed(1)
ob = (torch.rand(Rb, 3, generator=gb) * 2 - 1).to(dev); db = torch.randn(Rb, 3, generator=gb).to(dev)
zb = torch.sort(torch.rand(Rb, Sb, generator=gb) * 4 + 2, -1)[0].to(dev)
net = F.NerfNetwork.random(1, dev)
vdb, _ = F.ops.ray_setup(db)
raw = F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
torch.cuda.synchronize()
print("first launch ok", flush=True)
# odd sizes: tiles not a multiple of the grid, ragged last tile
r2 = F.ops.mlp_fwd(net.packed, ob[:5001], db[:5001], vdb[:5001], zb[:5001, :77], precision="bf16")
torch.cuda.synchronize()
for _ in range(10): raw = F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): raw = F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"cluster={os.environ.get('FNERF_MLP_CLUSTER', 'default')}: {ms:.3f} ms  {Rb*Sb*1186816/ms/1e9:7.1f} TFLOP/s  "
      f"checksum {raw.double().sum().item():.9e} / {r2.double().sum().item():.9e}", flush=True)
