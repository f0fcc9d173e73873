"""Round-robin A/B of fnerf_composite_fwd at R = 2^19 rays x S samples across libfnerf variants: python tools/ab_composite_fwd.py S lib..."""
import os, sys, torch, ctypes
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device("cuda:0")
S = int(sys.argv[1]); R = 1 << (19 if S <= 256 else 17)
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(R, S, 4, device=dev, generator=g)
z = torch.cumsum(torch.rand(R, S, device=dev, generator=g), -1) * (4.0 / S) + 2.0
dn = 1.0 + torch.rand(R, device=dev, generator=g)
rgb = torch.empty(R, 3, device=dev); dep = torch.empty(R, device=dev); acc = torch.empty(R, device=dev); dis = torch.empty(R, device=dev)
w = torch.empty(R, S, device=dev)
libs = []
for path in sys.argv[2:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fnerf_composite_fwd.restype = ctypes.c_int
    lib.fnerf_composite_fwd.argtypes = _lib.SIGNATURES["fnerf_composite_fwd"][1]
    libs.append((os.path.basename(path), lib))
def run(lib):
    rc = lib.fnerf_composite_fwd(raw.data_ptr(), z.data_ptr(), dn.data_ptr(), None, rgb.data_ptr(), dep.data_ptr(), acc.data_ptr(), dis.data_ptr(), w.data_ptr(), R, S, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
times = {n: [] for n, _ in libs}
for n, lib in libs:
    for _ in range(2): run(lib)
torch.cuda.synchronize()
for rnd in range(6):
    for n, lib in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(lib)
        e1.record(); torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / 5)
b = (24 * S + 36) * R
for n, tt in times.items():
    print(f"fwd S={S:4d} {n:24s} min {min(tt):.4f} ms  {b / min(tt) / 1e6:.0f} GB/s")
