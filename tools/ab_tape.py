"""A/B timing of the training forward (fnerf_mlp_fwd_tape) across libfnerf variants: python tools/ab_tape.py lib..."""
import os, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device('cuda:0')
Rb, Sb = 4096, 192
gb = torch.Generator().manual_seed(1)
ob = (torch.rand(Rb, 3, generator=gb) * 2 - 1).to(dev); db = torch.randn(Rb, 3, generator=gb).to(dev)
zb = torch.sort(torch.rand(Rb, Sb, generator=gb) * 4 + 2, -1)[0].to(dev)
for rep in range(2):
    for path in sys.argv[1:]:
        _lib._lib = None
        _lib.LIB_PATH = os.path.abspath(path)
        F.load_library()
        net = F.NerfNetwork.random(1, dev)
        vdb, _ = F.ops.ray_setup(db)
        lib = _lib.load()
        raw = torch.empty(Rb, Sb, 4, device=dev)
        tape = torch.empty(F.ops.mlp_tape_bytes(Rb, Sb), dtype=torch.uint8, device=dev)
        def run():
            _lib.check(lib.fnerf_mlp_fwd_tape(net.packed.data_ptr(), 0, ob.data_ptr(), db.data_ptr(), vdb.data_ptr(), zb.data_ptr(),
                                              None, None, 0, raw.data_ptr(), tape.data_ptr(), tape.numel(), Rb, Sb,
                                              torch.cuda.current_stream().cuda_stream), "fwd_tape")
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20): F.ops.mlp_fwd(net.packed, ob, db, vdb, zb, precision="bf16")
        e1.record(); torch.cuda.synchronize()
        print(f"{os.path.basename(path):36s} fwd_tape {ms:.3f} ms   plain fwd {e0.elapsed_time(e1) / 20:.3f} ms", flush=True)
