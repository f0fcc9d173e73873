"""Wait-cycle accounting of the pipelined backward (fnerf_debug_pipe_stats) on one 4096-ray training step."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F
from fashion_nerf_b200.train import Trainer
dev = torch.device("cuda:0")
lib = F.load_library()
model = F.NerfModel.random(dev)
o_all, d_all = F.pinhole_rays(800, 800)
idx = torch.randperm(800 * 800, generator=torch.Generator().manual_seed(0))[:4096]
o, d = o_all[idx].to(dev), d_all[idx].to(dev)
tgt = torch.rand(4096, 3).to(dev)
u_s, u_f = torch.rand(4096, 64).to(dev), torch.rand(4096, 128).to(dev)
tr = Trainer(model)
for _ in range(3):
    tr.step(o, d, tgt, 2.0, 6.0, 64, 128, u_strat=u_s, u_fine=u_f)
torch.cuda.synchronize()
n = lib.fnerf_debug_pipe_stats(None)
stats = torch.zeros(n, dtype=torch.int64, device=dev)
lib.fnerf_debug_pipe_stats(ctypes.c_void_p(stats.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.step(o, d, tgt, 2.0, 6.0, 64, 128, u_strat=u_s, u_fine=u_f)
e1.record()
torch.cuda.synchronize()
lib.fnerf_debug_pipe_stats(None)
print(f"step {e0.elapsed_time(e1):.3f} ms (coarse + fine backward accumulated below; 6 CTAs per role; Mcycles)")
names = ["V0a", "V0b", "V1a", "V1b", "F0", "F1"] + [f"L{l}_{h}" for l in range(7, 0, -1) for h in (0, 1)] + ["Z0a", "Z0b", "Z5a", "Z5b"]
hdr = ["ld:ring", "ld:slot", "mma:opnd", "mma:acc", "epi:acc", "epi:buf", "epi:stage", "st:img", "st:done", "st:read", "st:compl", "st:publ"]
print(f"{'role':6s}" + "".join(f"{h:>10s}" for h in hdr))
s = stats.view(-1, 12).cpu()
for r, nm in enumerate(names):
    print(f"{nm:6s}" + "".join(f"{s[r, k].item() / 6e6:10.2f}" for k in range(12)))
