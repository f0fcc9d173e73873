"""Debug driver: bf16 tcgen05 network query vs the fp32 kernel and the oracle (small case)."""
import sys, time, torch
sys.path.insert(0, '.')
from oracle import nerf_oracle as O
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
R, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
g = torch.Generator().manual_seed(12)
o = torch.rand(R, 3, generator=g) * 2 - 1
d = torch.randn(R, 3, generator=g)
z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0]
vd, _ = O.ray_setup(d)
p = O.init_params(0)
net = F.NerfNetwork.from_state_dict(p, dev)
args = [t.to(dev) for t in (o, d, vd, z)]
ref32 = F.ops.mlp_fwd(net.packed, *args, precision="fp32")
torch.cuda.synchronize()
print("fp32 kernel done", flush=True)
got = F.ops.mlp_fwd(net.packed, *args, precision="bf16")
torch.cuda.synchronize()
print("bf16 kernel done", flush=True)
err = (got - ref32).abs()
print("bf16 vs fp32 kernel: max abs", err.max().item(), "mean abs", err.mean().item(), "per-channel max", err.reshape(-1, 4).max(0)[0].tolist())
print("finite:", torch.isfinite(got).all().item(), "ref scale", ref32.abs().mean().item())
if R * S <= 65536:
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    emu = O.run_network(p, pts, vd, bf16=True)
    e2 = (got.cpu() - emu).abs()
    print("bf16 kernel vs oracle bf16-emulation: max abs", e2.max().item(), "mean", e2.mean().item())
    worst = err.reshape(-1, 4).max(1)[0].argmax().item()
    print("worst row", worst, got.reshape(-1, 4)[worst].tolist(), ref32.reshape(-1, 4)[worst].tolist())
    print("row0", got.reshape(-1, 4)[0].tolist(), ref32.reshape(-1, 4)[0].tolist())
# timing
for prec in ("bf16", "fp32"):
    Rb, Sb = 16384, 192
    gb = torch.Generator().manual_seed(1)
    ob = (torch.rand(Rb, 3, generator=gb) * 2 - 1).to(dev); db = torch.randn(Rb, 3, generator=gb).to(dev)
    zb = torch.sort(torch.rand(Rb, Sb, generator=gb) * 4 + 2, -1)[0].to(dev)
    vdb, _ = F.ops.ray_setup(db)
    for _ in range(10 if prec == "bf16" else 1): F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 40 if prec == "bf16" else 1
    for _ in range(n): F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision=prec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = Rb * Sb * 1186816
    print(f"{prec}: {ms:.3f} ms for {Rb*Sb} samples -> {fl/ms/1e9:.1f} TFLOP/s, {Rb*Sb/ms/1e3:.1f} Msamples/s")
