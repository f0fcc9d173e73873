"""torchrun worker for tests/test_dist_gpu.py: one fp32 training step per all-reduce mode on WORLD_SIZE GPUs, checked on
every rank against the oracle (mean of the per-rank oracle gradients -> oracle Adam)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fashion_nerf_b200 as F  # noqa: E402
from fashion_nerf_b200.train import Trainer  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Nc, Nf = 16, 16


def rank_data(r):
    o, d = O.pinhole_rays(8, 8, view=r, n_views=max(world, 2))
    g = torch.Generator().manual_seed(50 + r)
    return o, d, torch.rand(64, 3, generator=g), torch.rand(64, Nc, generator=g), torch.rand(64, Nf, generator=g)


# oracle: every rank's gradients on the CPU (all ranks compute all of them: tiny), mean, Adam
pc, pf = O.init_params(0), O.init_params(1)
for p in (pc, pf):
    p["alpha_linear.bias"] += 0.1
gc_sum, gf_sum = None, None
for r in range(world):
    o, d, tgt, u_s, u_f = rank_data(r)
    _, gc, gf = O.loss_and_grads(pc, pf, o, d, 2.0, 6.0, Nc, Nf, tgt, u_strat=u_s, u_fine=u_f)
    gc_sum = gc if gc_sum is None else {k: gc_sum[k] + gc[k] for k in gc}
    gf_sum = gf if gf_sum is None else {k: gf_sum[k] + gf[k] for k in gf}
ref_c = O.adam_step({k: v.clone() for k, v in pc.items()}, {k: v / world for k, v in gc_sum.items()}, {})
ref_f = O.adam_step({k: v.clone() for k, v in pf.items()}, {k: v / world for k, v in gf_sum.items()}, {})
ref_c, ref_f = F.flatten_state_dict(ref_c), F.flatten_state_dict(ref_f)
p0_c, p0_f = F.flatten_state_dict(pc), F.flatten_state_dict(pf)
gmean_c = F.flatten_state_dict({k: v / world for k, v in gc_sum.items()})
gmean_f = F.flatten_state_dict({k: v / world for k, v in gf_sum.items()})

o, d, tgt, u_s, u_f = (t.to(dev) for t in rank_data(rank))
ok = True
for mode in (False, True, "nvls"):
    model = F.NerfModel(F.NerfNetwork.from_state_dict(pc, dev), F.NerfNetwork.from_state_dict(pf, dev))
    try:
        tr = Trainer(model, fused_allreduce=mode)
    except Exception as e:                      # no multicast mapping on this box: NVLS mode unavailable
        if mode == "nvls":
            print(f"rank {rank}: nvls unavailable ({type(e).__name__})", flush=True)
            continue
        raise
    tr.step(o, d, tgt, 2.0, 6.0, Nc, Nf, u_strat=u_s, u_fine=u_f, precision="fp32")
    torch.cuda.synchronize()
    ec = (model.coarse.flat.cpu() - ref_c).abs().max().item()
    ef = (model.fine.flat.cpu() - ref_f).abs().max().item()
    # replicas identical?
    chk = torch.stack([model.coarse.flat.double().sum(), model.fine.flat.double().sum()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
    # Adam's first step moves every parameter by -lr * g / (|g| + eps) ~ -lr * sign(g).  A bound of ~2 lr on the parameters
    # would let a parameter step the WRONG way and pass, so: every parameter whose mean oracle gradient is above the fp32
    # path's rounding level (1e-3 of its tensor's largest gradient, at least 1e-6) must move against that gradient, and
    # the parameters must agree with the oracle's tightly on average.
    wrong = 0
    for net, ref_g, p0 in ((model.coarse, gmean_c, p0_c), (model.fine, gmean_f, p0_f)):
        delta = net.flat.cpu() - p0
        thr = torch.cat([torch.full((v.numel(),), max(1e-6, 1e-3 * v.abs().max().item())) for v in F.unflatten(ref_g).values()])
        sure = ref_g.abs() > thr
        wrong += int((torch.sign(delta[sure]) != -torch.sign(ref_g[sure])).sum())
    mean_c = (model.coarse.flat.cpu() - ref_c).abs().mean().item()
    mean_f = (model.fine.flat.cpu() - ref_f).abs().mean().item()
    good = same and wrong == 0 and mean_c <= 2e-5 and mean_f <= 2e-5
    ok = ok and good
    print(f"rank {rank} mode {mode}: wrong-way steps {wrong}, max|dparam| coarse {ec:.2e} fine {ef:.2e} mean {mean_c:.2e} / {mean_f:.2e} "
          f"replicas_identical {same}", flush=True)
dist.barrier()
dist.destroy_process_group()
print("DIST_OK" if ok else "DIST_FAIL", flush=True)
sys.exit(0 if ok else 1)
