"""BASELINE.json configs[3] and configs[4] (and the fp32 variant of configs[1]) through the public API.

  --config 4 : 1920x1080 frame, 256 coarse + 768 fine samples, ray-sharded over WORLD_SIZE ranks
               (one contiguous slice of the flattened ray list per rank, no collective)
  --config 5 : garment-latent-conditioned NeRF, 32 views of 512x512, views sharded over ranks,
               one 256-d code per view
  --config 2f: 800x800, 64+128, fp32 SIMT path
Run under torchrun for N > 1, or with --emulate-world N on one GPU to time rank 0's shard of an N-GPU
job (render has no communication, so the shard time IS the N-GPU step time).  One JSON line on rank 0."""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fashion_nerf_b200 as F

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="4")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--emulate-world", type=int, default=0)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
eff_world = args.emulate_world or world

if args.config == "4":
    H, W, Nc, Nf, prec, cond = 1080, 1920, 256, 768, "bf16", False
    o, d = F.pinhole_rays(H, W)
    R_total = o.shape[0]
    lo, hi = R_total * rank // eff_world, R_total * (rank + 1) // eff_world
    o, d = o[lo:hi].to(dev), d[lo:hi].to(dev)
    view_id = codes = None
    flop_per_ray = 1_186_816 * (Nc + Nc + Nf)
elif args.config == "5":
    H = W = 512; Nc, Nf, prec, cond, V = 64, 128, "bf16", True, 32
    views = [v for v in range(V) if v % eff_world == rank]
    os_, ds_, vid = [], [], []
    for v in views:
        a, b = F.pinhole_rays(H, W, view=v, n_views=V)
        os_.append(a); ds_.append(b); vid.append(torch.full((H * W,), v, dtype=torch.int32))
    o, d, view_id = torch.cat(os_).to(dev), torch.cat(ds_).to(dev), torch.cat(vid).to(dev)
    codes = torch.randn(V, 256, generator=torch.Generator().manual_seed(2)).to(dev)
    R_total = V * H * W
    flop_per_ray = 1_186_816 * (Nc + Nc + Nf)      # un-conditioned figure (cond term is hoisted per view)
else:
    H = W = 800; Nc, Nf, prec, cond = 64, 128, "fp32", False
    o, d = F.pinhole_rays(H, W)
    R_total = o.shape[0]
    lo, hi = R_total * rank // eff_world, R_total * (rank + 1) // eff_world
    o, d = o[lo:hi].to(dev), d[lo:hi].to(dev)
    view_id = codes = None
    flop_per_ray = 1_186_816 * (Nc + Nc + Nf)

model = F.NerfModel.random(dev, cond=cond)
R = o.shape[0]
chunk = 1 << 18 if args.config == "4" else 1 << 20


def step():
    return F.render_image(model, o, d, 2.0, 6.0, Nc, Nf, codes, chunk=chunk, precision=prec,
                          **({"view_id": view_id} if view_id is not None else {}))


for _ in range(args.warmup):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
if rank == 0:
    print(json.dumps({"config": args.config, "world": eff_world, "emulated": bool(args.emulate_world), "rays_this_rank": R,
                      "rays_total": R_total, "N_samples": Nc, "N_importance": Nf, "precision": prec, "ms_per_step": ms,
                      "Mrays_per_s_job": R_total / ms / 1e3 if eff_world == world or args.emulate_world else None,
                      "Mrays_per_s_per_gpu": R / ms / 1e3, "TFLOPs_per_gpu": R * flop_per_ray / ms / 1e9,
                      "rgb_finite": bool(torch.isfinite(out["rgb"]).all())}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
