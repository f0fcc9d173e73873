"""Round-robin A/B of fnerf_composite_bwd at R = 2^18 rays x S samples across libfnerf variants: python tools/ab_composite_bwd.py S lib..."""
import os, sys, torch, ctypes, statistics
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device("cuda:0")
S = int(sys.argv[1]); R = 1 << 18
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(R, S, 4, device=dev, generator=g)
z = torch.cumsum(torch.rand(R, S, device=dev, generator=g), -1) * (4.0 / S) + 2.0
dn = 1.0 + torch.rand(R, device=dev, generator=g)
g_rgb = torch.randn(R, 3, device=dev, generator=g)
out = torch.empty(R, S, 4, device=dev)
libs = []
for path in sys.argv[2:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fnerf_composite_bwd.restype = ctypes.c_int
    lib.fnerf_composite_bwd.argtypes = _lib.SIGNATURES["fnerf_composite_bwd"][1]
    libs.append((os.path.basename(path), lib))
def run(lib):
    rc = lib.fnerf_composite_bwd(raw.data_ptr(), z.data_ptr(), dn.data_ptr(), None, g_rgb.data_ptr(), None, None, out.data_ptr(), R, S, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
times = {n: [] for n, _ in libs}
for n, lib in libs:
    for _ in range(2): run(lib)
torch.cuda.synchronize()
for rnd in range(6):
    for n, lib in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): run(lib)
        e1.record(); torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / 3)
b = (36 * S + 24) * R
for n, tt in times.items():
    print(f"{n:28s} min {min(tt):.3f} ms  {b / min(tt) / 1e6:.0f} GB/s")
