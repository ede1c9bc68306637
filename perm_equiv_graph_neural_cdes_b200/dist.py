"""Multi-GPU data parallelism for the hot path: the batch of graph trajectories is sharded over the
ranks (the reference's ``jax.vmap(model)`` batch, src/configs/loss_configs.py:44 /
src/engine/trainer_oversampling.py:47), every rank runs the fused solve + adjoint on its shard with no
data-path collective, and the small flat parameter-gradient buffer (0.7k .. 150k floats) is summed with
ONE all-reduce per step (NCCL over NVLink on GPUs; gloo in the CPU tests).

One process per GPU, launched with torchrun.  Nothing here touches the kernels.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``total`` trajectories for ``rank`` (first ranks get the remainder)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int, dim: int = 0) -> List[torch.Tensor]:
    """Slices every tensor's batch dimension to this rank's trajectories."""
    out = []
    for t in tensors:
        b, e = shard_range(t.shape[dim], rank, world)
        out.append(t.narrow(dim, b, e - b))
    return out


def allreduce_gradients(params: Sequence[torch.nn.Parameter], average_over: int = 0, group=None) -> torch.Tensor:
    """Sums (optionally averages) the gradients of ``params`` across ranks through ONE flat buffer and writes them
    back in place.  Returns the reduced flat buffer (what optax's clip_by_global_norm / adamw consume in the
    reference, src/configs/optimiser_configs.py:70-88)."""
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average_over:
        flat = flat / float(average_over)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[off:off + n].view_as(g).clone()
        off += n
    return flat


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
