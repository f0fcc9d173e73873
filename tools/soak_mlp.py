"""Soak: many launches of the render kernel (mode from FNERF_MLP_CLUSTER) and of the tape forward / backward on
varying sizes; every repeat of a size must reproduce its first checksum bit for bit (catches rare ordering bugs)."""
import os, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(3)
net = F.NerfNetwork.random(2, dev)
sizes = [(16384, 192), (4096, 64), (5001, 77), (333, 192), (65536, 64), (1, 1), (129, 1), (20000, 129)]
data = {}
for R, S in sizes:
    o = (torch.rand(R, 3, generator=g) * 2 - 1).to(dev); d = torch.randn(R, 3, generator=g).to(dev)
    z = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, -1)[0].to(dev)
    vd, _ = F.ops.ray_setup(d)
    data[(R, S)] = (o, d, vd, z, torch.randn(R, S, 4, generator=g).to(dev))
first = {}
n_launch = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    for key, (o, d, vd, z, graw) in data.items():
        raw = F.ops.mlp_fwd(net.packed, o, d, vd, z, precision="bf16")
        chk = [raw.view(torch.int32).to(torch.int64).sum().item()]
        if rep % 8 == 0:
            raw2, tape = F.ops.mlp_fwd_tape(net.packed, o, d, vd, z)
            fg = torch.zeros(net.flat.numel(), device=dev)
            F.ops.mlp_bwd_tape(net.packed, graw, tape, fg)
            chk += [raw2.view(torch.int32).to(torch.int64).sum().item(), bool(torch.isfinite(fg).all())]
            n_launch += 3
        n_launch += 1
        k = (key, len(chk))
        if k not in first: first[k] = chk
        assert first[k] == chk, (key, rep, first[k], chk)
torch.cuda.synchronize()
print(f"soak ok: {n_launch} launches, mode {os.environ.get('FNERF_MLP_CLUSTER', 'default')}")
