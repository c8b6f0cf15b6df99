"""Multi-GPU plumbing: one process per GPU (torchrun), envs sharded by contiguous global env-id
ranges with no data-path collective; the only collective is the DQN gradient all-reduce
(SURVEY.md 8e).  Pure torch.distributed: works with NCCL on GPUs and with gloo on CPU (tests)."""
from __future__ import annotations

import os
from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_env_ids(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous global env-id range [first, first + count) owned by `rank`.  The per-env RNG is
    keyed by the GLOBAL id, so a trajectory does not depend on how many GPUs share the work."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(total_envs, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


@torch.no_grad()
def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], world: int, assign_views: bool = False) -> None:
    """Average gradients across ranks with ONE all-reduce of the flattened gradient
    (156 934 floats = 628 KB for QNetV2: latency-bound, so a single bucket).  The gradients are
    flattened in MEMORY order (the reduction is elementwise and every rank holds the same layouts),
    so with `assign_views=True` each `p.grad` simply becomes a view of the reduced flat buffer with
    the strides it had (NHWC weights keep NHWC gradients): one `cat`, the collective and one
    `div_` instead of those plus a copy back per parameter."""
    if world <= 1:
        return
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    grads = [p.grad for p in params]
    dense = all(g.is_contiguous() or (g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last))
                for g in grads)
    raw = [g.as_strided((g.numel(),), (1,)) if dense else g.reshape(-1) for g in grads]
    flat = torch.cat(raw)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    offset = 0
    for p, g, r in zip(params, grads, raw):
        if assign_views and dense:
            p.grad = flat.as_strided(g.size(), g.stride(), offset)
        elif dense:
            r.copy_(flat[offset:offset + g.numel()])
        else:
            g.copy_(flat[offset:offset + g.numel()].view_as(g))
        offset += g.numel()


def max_over_ranks(value: float, device: torch.device, world: int) -> float:
    """Timing reduction of bench.py: the slowest rank defines the step time."""
    if world <= 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
