"""Batch-sharded sampling across the GPUs of one box.

The reference has no distributed code.  The sampling path shards along the batch:
rank r runs the whole step loop on its own slice (its own prior seed, its own
E2-CRF cache keyed on its local sample 0 -- exactly the reference's semantics of
``world`` independent runs with ``num_samples = sample_batch_size = B/world``),
with no per-step collective; the only communication is one all-gather of the
final ``[B/world, L, C]`` fp32 samples (NCCL over NVLink on GPUs, gloo in the
CPU tests of this host logic).
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_samples: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of rank ``rank``; the first ``num_samples % world_size``
    ranks take one extra sample, so ragged counts are covered exactly once."""
    assert 0 <= rank < world_size and num_samples >= 0
    base, extra = divmod(num_samples, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_samples(local: torch.Tensor, num_samples: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather the per-rank samples into ``[num_samples, L, C]`` on every rank (rank order = sample order).
    Shards may differ by one row, so each rank pads to the largest shard and the padding is dropped."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(num_samples, world, r) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: e - b] for p, (b, e) in zip(parts, sizes)], dim=0)


def sample_sharded(sample_fn: Callable[[int, int], torch.Tensor], num_samples: int,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Run ``sample_fn(n_local, rank)`` (e.g. ``lambda n, r: sampler.run(...)``) on this rank's slice and
    gather.  ``sample_fn`` must return ``[n_local, L, C]`` on the device the process group communicates on."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    b, e = shard_bounds(num_samples, world, rank)
    local = sample_fn(e - b, rank)
    assert local.shape[0] == e - b
    return gather_samples(local, num_samples, group)
