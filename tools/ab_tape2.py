"""Round-robin A/B timing of the training forward (fnerf_mlp_fwd_tape) across libfnerf variants, all loaded at once:
   python tools/ab_tape2.py lib...   -> min / median ms per variant over 12 rounds of 10 launches (4096 rays x 192 samples)."""
import ctypes, os, statistics, sys, torch
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device('cuda:0')
Rb, Sb = 4096, 192
gb = torch.Generator().manual_seed(1)
ob = (torch.rand(Rb, 3, generator=gb) * 2 - 1).to(dev); db = torch.randn(Rb, 3, generator=gb).to(dev)
zb = torch.sort(torch.rand(Rb, Sb, generator=gb) * 4 + 2, -1)[0].to(dev)
F.load_library()
net = F.NerfNetwork.random(1, dev)
vdb, _ = F.ops.ray_setup(db)
raw = torch.empty(Rb, Sb, 4, device=dev)
tape = torch.empty(F.ops.mlp_tape_bytes(Rb, Sb), dtype=torch.uint8, device=dev)
libs = []
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fnerf_mlp_fwd_tape.restype = ctypes.c_int
    lib.fnerf_mlp_fwd_tape.argtypes = _lib.SIGNATURES["fnerf_mlp_fwd_tape"][1]
    libs.append((os.path.basename(path), lib))
st = torch.cuda.current_stream().cuda_stream
c = ctypes.c_void_p
def run(lib):
    rc = lib.fnerf_mlp_fwd_tape(net.packed.data_ptr(), 0, ob.data_ptr(), db.data_ptr(), vdb.data_ptr(), zb.data_ptr(),
                                None, None, 0, raw.data_ptr(), tape.data_ptr(), tape.numel(), Rb, Sb, st)
    assert rc == 0, rc
times = {n: [] for n, _ in libs}
for n, lib in libs:
    for _ in range(5): run(lib)
torch.cuda.synchronize()
for rnd in range(12):
    for n, lib in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run(lib)
        e1.record(); torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / 10)
for n, t in times.items():
    print(f"{n:36s} min {min(t):.3f}  median {statistics.median(t):.3f} ms", flush=True)
