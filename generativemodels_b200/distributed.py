"""Batch-of-samples sharding across the GPUs of one box.

Every sample's trajectory is independent, so the path shards with no data-path collective: rank r of R takes a
contiguous slice of the batch, runs the unchanged single-GPU sampler on it, and the finished samples are gathered
once at the end (``all_gather`` over NCCL/NVLink; 22.9 MB per 160x224x160 fp32 volume).  Noise is drawn once with the
global seed on the host and sliced per rank, so results do not depend on R.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, near-equal split: the first ``n % world`` ranks get one extra sample."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_samples(local: torch.Tensor | None, total: int, device=None) -> torch.Tensor:
    """All ranks receive the full batch ``[total, ...]`` in rank order (ragged shards are padded).  A rank whose shard
    is empty (more ranks than samples) passes ``None``: it learns the per-sample shape from the others and still takes
    part in the collective, so nobody is left blocked in ``all_gather``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = max(shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world))
    if total < world:                      # only then can a shard be empty: agree on the per-sample shape first
        metas = [None] * world
        dist.all_gather_object(metas, None if local is None else (tuple(local.shape[1:]), local.dtype))
        tail, dtype = next(m for m in metas if m is not None)
        if local is None:
            dev = device if device is not None else torch.device("cuda", torch.cuda.current_device()) \
                if dist.get_backend() == "nccl" else torch.device("cpu")
            local = torch.zeros((0, *tail), dtype=dtype, device=dev)
    pad = local
    if local.shape[0] < per:
        pad = torch.cat([local, local.new_zeros((per - local.shape[0], *local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    parts = []
    for r, b in enumerate(bufs):
        lo, hi = shard_bounds(total, r, world)
        parts.append(b[: hi - lo])
    return torch.cat(parts, 0)


def sample_sharded(sample_fn, input_noise: torch.Tensor, **kwargs) -> torch.Tensor:
    """Run ``sample_fn(input_noise=shard, **kwargs)`` on this rank's shard of the batch and gather the result.
    Batch-indexed keyword tensors (``conditioning``, ``cn_cond``) are sharded alongside the noise."""
    total = input_noise.shape[0]
    local_kwargs = {k: (shard_batch(v) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == total else v)
                    for k, v in kwargs.items()}
    shard = shard_batch(input_noise)
    out = sample_fn(input_noise=shard, **local_kwargs) if shard.shape[0] > 0 else None
    return gather_samples(out, total, device=input_noise.device)
