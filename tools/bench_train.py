"""BASELINE.json configs[2]: data-parallel training step, 4096 rays/GPU, fwd+bwd through compositing and
the MLP, NCCL all-reduce of the flat 4.77 MB gradient buffer, Adam, re-pack.  Launch with torchrun for
N > 1 (one rank per GPU).  Prints one JSON line (rank 0): ms/step (device-timed, max over ranks),
Mrays/s, and the all-reduce's share measured with CUDA events."""
import argparse, json, os, statistics, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F
from fashion_nerf_b200.train import Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--allreduce", default="auto", choices=["auto", "fused", "nvls", "nccl"])
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model = F.NerfModel.random(dev)
o_all, d_all = F.pinhole_rays(800, 800)
idx = torch.randperm(800 * 800, generator=torch.Generator().manual_seed(rank))[:4096]   # seed = rank
o, d = o_all[idx].to(dev), d_all[idx].to(dev)
tgt = torch.rand(4096, 3, generator=torch.Generator().manual_seed(100 + rank)).to(dev)
g = torch.Generator().manual_seed(rank)
u_s, u_f = torch.rand(4096, 64, generator=g).to(dev), torch.rand(4096, 128, generator=g).to(dev)
tr = Trainer(model, fused_allreduce={"auto": None, "fused": True, "nvls": "nvls", "nccl": False}[args.allreduce])
times, losses = [], []
for i in range(args.warmup + args.steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = tr.step(o, d, tgt, 2.0, 6.0, 64, 128, u_strat=u_s, u_fine=u_f, precision=args.precision)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if i >= args.warmup:
        times.append(ms)
    losses.append(res["loss"].item())
# replicas must stay identical
if world > 1:
    chk = model.coarse.flat.double().sum().reshape(1)
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
else:
    same = True
if rank == 0:
    ms = statistics.mean(times)
    flop = 4096 * 893_190_144            # train fwd+bwd FLOP/ray
    print(json.dumps({"bench": "train_step", "n_gpus": world, "rays_per_gpu": 4096, "ms_per_step": ms,
                      "Mrays_per_s": world * 4096 / ms / 1e3, "algorithmic_TFLOPs_per_gpu": flop / ms / 1e9,
                      "precision_fwd": args.precision, "bwd": "bf16 tcgen05 (tape + dgrad + wgrad)" if args.precision == "bf16" else "fp32 SGEMM chain", "loss_first_last": [losses[0], losses[-1]],
                      "replicas_identical": same,
                      "allreduce": ("NVLS multimem all-reduce + Adam" if tr.nvls else "fused P2P all-reduce + Adam (symmetric memory)") if tr.symm is not None else ("NCCL" if world > 1 else "none"),
                      "param_checksum": model.coarse.flat.double().sum().item()}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
