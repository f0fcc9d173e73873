"""world_size=2 gloo tests (CPU) of the host-side data-parallel logic: the flat gradient all-reduce and
Adam keep replicas identical and equal to a single-process run on the concatenated batch; ray shards
partition the frame without overlap."""
import os
import socket

import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _cpu_adam import cpu_adam_update  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fashion_nerf_b200.train import FlatAdam, allreduce_mean_
    n = 1_191_688                                                     # coarse + fine flat parameters
    g = torch.Generator().manual_seed(7)
    params = torch.randn(n, generator=g)
    opt = FlatAdam(n, "cpu", update_fn=cpu_adam_update)
    for step in range(3):
        # per-rank "local" gradient of a quadratic loss on this rank's shard of a synthetic batch
        data = torch.randn(n, generator=torch.Generator().manual_seed(100 * step + rank))
        grad = (params - data)
        allreduce_mean_(grad)
        opt.step(params, grad)
    gathered = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(gathered, params)
    if rank == 0:
        out.put([t.clone() for t in gathered])
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_adam_two_ranks_match_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    replicas = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(replicas[0], replicas[1])                      # replicas stay bit-identical
    # single-process reference: mean of the two local gradients each step
    from fashion_nerf_b200.train import FlatAdam
    n = 1_191_688
    params = torch.randn(n, generator=torch.Generator().manual_seed(7))
    opt = FlatAdam(n, "cpu", update_fn=cpu_adam_update)
    for step in range(3):
        grads = [params - torch.randn(n, generator=torch.Generator().manual_seed(100 * step + r)) for r in range(world)]
        opt.step(params, (grads[0] + grads[1]) / world)
    assert torch.allclose(replicas[0], params, atol=1e-6)


def test_flat_adam_matches_torch_optim():
    from fashion_nerf_b200.train import FlatAdam
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g)
    a = p0.clone()
    b = p0.clone().requires_grad_(True)
    mine, ref = FlatAdam(1000, "cpu", lr=5e-4, update_fn=cpu_adam_update), torch.optim.Adam([b], lr=5e-4)
    for _ in range(5):
        grad = torch.randn(1000, generator=g)
        mine.step(a, grad)
        b.grad = grad.clone()
        ref.step()
    assert torch.allclose(a, b.detach(), atol=1e-7)


def test_flat_adam_has_no_cpu_arithmetic_of_its_own():
    """Without an injected update function the optimiser goes to the CUDA kernel and refuses CPU buffers."""
    import pytest
    from fashion_nerf_b200 import FnerfError
    from fashion_nerf_b200.train import FlatAdam
    opt = FlatAdam(8, "cpu")
    with pytest.raises(FnerfError):
        opt.step(torch.zeros(8), torch.ones(8))


def test_ray_shards_partition_the_frame():
    """Contiguous 1/P slices of the flattened ray list: disjoint, complete, balanced to one ray."""
    R = 800 * 800
    for P in (1, 2, 4, 8, 3):
        bounds = [(R * r // P, R * (r + 1) // P) for r in range(P)]
        assert bounds[0][0] == 0 and bounds[-1][1] == R
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(P - 1))
        sizes = [b - a for a, b in bounds]
        assert max(sizes) - min(sizes) <= 1
