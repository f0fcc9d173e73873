"""Round-robin A/B of fnerf_stratified at R = 2^20 rays, N = 64 / 256 across libfnerf variants: python tools/ab_stratified.py lib..."""
import os, sys, torch, ctypes
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device("cuda:0")
R = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
near, far = torch.full((R,), 2.0, device=dev), torch.full((R,), 6.0, device=dev)
libs = []
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fnerf_stratified.restype = ctypes.c_int
    lib.fnerf_stratified.argtypes = _lib.SIGNATURES["fnerf_stratified"][1]
    libs.append((os.path.basename(path), lib))
for N in (64, 256):
    t = torch.linspace(0, 1, N).to(dev)
    u = torch.rand(R, N, device=dev, generator=g)
    z = torch.empty(R, N, device=dev)
    def run(lib):
        rc = lib.fnerf_stratified(near.data_ptr(), far.data_ptr(), t.data_ptr(), u.data_ptr(), z.data_ptr(), R, N, 0, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    times = {n: [] for n, _ in libs}
    for n, lib in libs:
        for _ in range(3): run(lib)
    torch.cuda.synchronize()
    for rnd in range(8):
        for n, lib in libs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run(lib)
            e1.record(); torch.cuda.synchronize()
            times[n].append(e0.elapsed_time(e1) / 10)
    for n, tt in times.items():
        print(f"N={N:4d} {n:26s} min {min(tt):.4f} ms  {(8 * N + 8) * R / min(tt) / 1e6:.0f} GB/s")
