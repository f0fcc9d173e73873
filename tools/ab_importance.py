import os, sys, torch, ctypes, statistics
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
from fashion_nerf_b200 import _lib
dev = torch.device("cuda:0")
R = 1 << 20
g = torch.Generator(device=dev).manual_seed(0)
near, far = torch.full((R,), 2.0, device=dev), torch.full((R,), 6.0, device=dev)
t = torch.linspace(0, 1, 64).to(dev)
zc = F.ops.stratified(near, far, t, torch.rand(R, 64, device=dev, generator=g))
w = torch.rand(R, 64, device=dev, generator=g)
u = torch.rand(R, 128, device=dev, generator=g)
zs = torch.empty(R, 128, device=dev); zf = torch.empty(R, 192, device=dev); zstd = torch.empty(R, device=dev)
libs = []
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fnerf_importance.restype = ctypes.c_int
    lib.fnerf_importance.argtypes = _lib.SIGNATURES["fnerf_importance"][1]
    libs.append((os.path.basename(path), lib))
def run(lib):
    rc = lib.fnerf_importance(zc.data_ptr(), w.data_ptr(), u.data_ptr(), 128, zs.data_ptr(), zf.data_ptr(), None, zstd.data_ptr(), R, 64, 128, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
times = {n: [] for n, _ in libs}
for n, lib in libs:
    for _ in range(3): run(lib)
torch.cuda.synchronize()
for rnd in range(8):
    for n, lib in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(lib)
        e1.record(); torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / 5)
for n, tt in times.items():
    print(f"{n:30s} min {min(tt):.4f} median {statistics.median(tt):.4f} ms")
