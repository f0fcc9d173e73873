"""A/B timing of libfnerf variants in ONE process/box: python tools/ab_tc.py libA.so libB.so ..."""
import ctypes, os, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device('cuda:0')
Rb, Sb = 16384, 192
gb = torch.Generator().manual_seed(1)
ob = (torch.rand(Rb, 3, generator=gb) * 2 - 1).to(dev); db = torch.randn(Rb, 3, generator=gb).to(dev)
zb = torch.sort(torch.rand(Rb, Sb, generator=gb) * 4 + 2, -1)[0].to(dev)
ref = None
for rep in range(2):
    for path in sys.argv[1:]:
        _lib._lib = None
        _lib.LIB_PATH = os.path.abspath(path)
        F.load_library()
        net = F.NerfNetwork.random(1, dev)
        vdb, _ = F.ops.ray_setup(db)
        for _ in range(10): raw = F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): raw = F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        if ref is None: ref = raw.clone()
        print(f"{os.path.basename(path):32s} {ms:.3f} ms  {Rb*Sb*1186816/ms/1e9:7.1f} TFLOP/s  maxdiff vs first {(raw-ref).abs().max().item():.2e}", flush=True)
