"""Slice sharding across the GPUs of one node: every slice's trajectory is independent, so rank r of W
takes a contiguous range of the slice batch, weights are replicated, and the ONLY collective is a gather of the final
latents (NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_slices: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the slices rank ``rank`` owns: contiguous, sizes differ by at most one, earlier ranks take the extra."""
    if not (0 <= rank < world) or n_slices < 0:
        raise ValueError("bad rank/world/n_slices")
    base, rem = divmod(n_slices, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_slices(local: torch.Tensor, n_slices: int) -> torch.Tensor:
    """All-gather per-rank results ``[n_local, ...]`` into ``[n_slices, ...]`` in global slice order (ragged tails are
    padded to the largest shard for the collective and trimmed afterwards)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_slices, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def sharded_apply(n_slices: int, process: Callable[[int, int], torch.Tensor]) -> torch.Tensor:
    """BASELINE config 5 (strong scaling): the ``n_slices`` slices of a sweep are split into contiguous per-rank ranges
    (``shard_range``), every rank runs ``process(lo, hi) -> [hi - lo, ...]`` on its own range with NO communication, and ONE
    ``all_gather`` at the end returns all results in global slice order on every rank.  Single process: ``process(0, n)``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return process(0, n_slices)
    lo, hi = shard_range(n_slices, dist.get_rank(), dist.get_world_size())
    return gather_slices(process(lo, hi), n_slices)
