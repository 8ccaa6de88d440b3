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


def warm_communicator(device: torch.device, group: Optional[dist.ProcessGroup] = None) -> None:
    """One tiny all-gather before anything is timed.  NCCL connects its channels lazily, per collective algorithm, at
    the first call that needs them: the first all-gather of an 8-rank communicator spends ~50 ms there (that, not
    the 12 MB payload, was the 53.6 ms "final gather" of the round-1 scaling run; at 2 and 4 ranks the set-up is
    shorter and was hidden by the preceding barrier's all-reduce)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    world = dist.get_world_size(group)
    src = torch.zeros(256, device=device)
    dst = torch.empty(256 * world, device=device)
    dist.all_gather_into_tensor(dst, src, group=group)
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def gather_samples(local: torch.Tensor, num_samples: int, group: Optional[dist.ProcessGroup] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather the per-rank samples into ``[num_samples, L, C]`` on every rank (rank order = sample order): ONE
    ``all_gather_into_tensor`` straight into the (optionally preallocated) result.  Shards differ by at most one row
    (``shard_bounds``); only then each rank pads to the largest shard and the padding rows are dropped afterwards."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(num_samples, world, r) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    tail = tuple(local.shape[1:])
    local = local.contiguous()
    if num_samples % world == 0:
        if out is None:
            out = torch.empty((num_samples,) + tail, dtype=local.dtype, device=local.device)
        assert tuple(out.shape) == (num_samples,) + tail and out.is_contiguous()
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    pad = torch.zeros((biggest,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * biggest,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    parts = [buf[r * biggest: r * biggest + (e - b)] for r, (b, e) in enumerate(sizes)]
    if out is None:
        return torch.cat(parts, dim=0)
    torch.cat(parts, dim=0, out=out)
    return out


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
