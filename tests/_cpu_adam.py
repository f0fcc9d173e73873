"""Test-side stand-in for the fused CUDA Adam kernel (fnerf_adam_step), so the world_size>1 host logic of
fashion_nerf_b200.train (all-reduce, step bookkeeping, slicing of the flat buffers) can be exercised on CPU with gloo.
It is injected into FlatAdam(update_fn=...); the product package holds no CPU arithmetic."""
import torch


def cpu_adam_update(params, grad, m, v, t, lr, b1, b2, eps, grad_scale):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) on CPU tensors, in place."""
    if grad_scale != 1.0:
        grad = grad * grad_scale
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = (v / bc2).sqrt_().add_(eps)
    params.addcdiv_(m, denom, value=-lr / bc1)
