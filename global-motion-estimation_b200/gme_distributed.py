"""Frame-pair sharding across the GPUs of one box (one process per GPU, torch.distributed).

Every frame pair (k, k + d) is independent (results.py:41-50), so the sequence is partitioned
into contiguous pair ranges with NO cross-GPU traffic on the hot path; the only collective is
one final all-gather of [pairs, 7] float64 rows (6 affine parameters + PSNR), 56 bytes per pair.
The compute step is injected, so the partition/gather logic is testable on CPU with gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_pairs(n_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition: rank r owns pairs [start, stop).  Contiguity keeps the frames a
    rank needs to [start, stop + d), so neighbouring pairs share frames (and cached pyramids)."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local: torch.Tensor, n_pairs: int, group=None) -> torch.Tensor:
    """local: float64[n_local, k] rows of this rank's pair range -> float64[n_pairs, k] on every rank."""
    world = dist.get_world_size(group)
    k = local.shape[1]
    if n_pairs % world == 0 and dist.get_backend(group) == "nccl" and local.is_contiguous():
        # equal shards (the bench, and any sequence whose pair count divides): one collective, no padding, no copies
        out = torch.empty((n_pairs, k), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    width = (n_pairs + world - 1) // world                # longest shard; shorter shards are padded
    padded = torch.zeros((width, k), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * width, k), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, padded, group=group)
    else:                                                 # gloo (CPU tests)
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        out = torch.cat(parts, 0)
    rows = []
    for r in range(world):
        a, b = shard_pairs(n_pairs, r, world)
        rows.append(out[r * width:r * width + (b - a)])
    return torch.cat(rows, 0)


def gather_rows_async(local: torch.Tensor, out: torch.Tensor, group=None):
    """Equal shards, NCCL: starts ONE all_gather_into_tensor of this rank's contiguous [n_local, k] rows into
    ``out`` ([world * n_local, k]) on NCCL's own stream and returns the work handle.  The collective waits for the
    kernels already enqueued on the current stream, but nothing enqueued afterwards waits for it: the next batch's
    kernels run while the rows travel.  ``work.wait()`` orders the current stream after the gather."""
    return dist.all_gather_into_tensor(out, local, group=group, async_op=True)


def run_sharded(n_pairs: int, compute, group=None) -> torch.Tensor:
    """compute(start, stop) -> float64[stop - start, k] for this rank's pairs; returns all rows on all ranks."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    start, stop = shard_pairs(n_pairs, rank, world)
    return gather_rows(compute(start, stop), n_pairs, group)
