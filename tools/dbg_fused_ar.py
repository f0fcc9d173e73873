"""2-GPU timing of the pieces of the fused all-reduce + Adam step vs NCCL all-reduce + Adam (torchrun)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, '.')
import fashion_nerf_b200 as F
import torch.distributed._symmetric_memory as symm_mem
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1_191_688
buf = symm_mem.empty(n, dtype=torch.float32, device=dev); h = symm_mem.rendezvous(buf, dist.group.WORLD)
buf.normal_()
p = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
g2 = torch.randn(n, device=dev)
def timeit(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
t_bar = timeit(lambda: h.barrier(channel=0))
t_k = timeit(lambda: F.ops.allreduce_adam_step(h.buffer_ptrs_dev, world, 0, p, m, v, 1))
t_all = timeit(lambda: (h.barrier(channel=0), F.ops.allreduce_adam_step(h.buffer_ptrs_dev, world, 0, p, m, v, 1), h.barrier(channel=1)))
t_nccl = timeit(lambda: (dist.all_reduce(g2), g2.div_(world), F.ops.adam_step(p, g2, m, v, 1)))
# NVLS two-shot: correctness against NCCL, then timing
t_mm = -1.0
mc = h.multicast_ptr
if mc:
    n4 = n // 4 * 4
    buf.copy_(torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank)))
    ref = buf.clone(); dist.all_reduce(ref)
    h.barrier(channel=0)
    F.ops.multimem_allreduce(mc, rank, world, n4, dev)
    h.barrier(channel=1)
    torch.cuda.synchronize()
    err = (buf[:n4] - ref[:n4]).abs().max().item()
    chk = buf[:n4].double().sum().reshape(1); allc = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(allc, chk)
    same = all(torch.equal(allc[0], c) for c in allc)
    t_mm = timeit(lambda: (h.barrier(channel=0), F.ops.multimem_allreduce(mc, rank, world, n4, dev), h.barrier(channel=1),
                           F.ops.adam_step(p, buf, m, v, 1, grad_scale=1.0 / world)))
    if rank == 0: print(f"multimem: max |err| vs NCCL {err:.3e}, identical on all ranks {same}, barrier+NVLS+barrier+adam {t_mm:.1f} us")
if rank == 0:
    print(f"world {world}: symm barrier {t_bar:.1f} us | fused kernel {t_k:.1f} us | barrier+kernel+barrier {t_all:.1f} us | NCCL all_reduce + div + adam {t_nccl:.1f} us | multicast ptr {hex(mc) if mc else 0}")
dist.destroy_process_group()
