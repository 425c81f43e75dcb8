"""Ray / view data parallelism for the render path: one process per GPU, the grid replicated, rays sharded, and ONE
all-reduce of the voxel gradients per optimiser step (the reference itself is single-GPU).

Rays are independent in the forward pass, so there is no exchange there.  The backward pass leaves every rank with a
partial dense gradient of the (replicated) grid; summing those partials is the only collective of a training step:
``all_reduce(SUM)`` over one flat fp32 buffer of (F+1)*X*Y*Z elements (65.5 MB at 160^3 SH-0), issued on the compute
stream right after the backward kernel.  ``torch.distributed`` is the plumbing (NCCL over NVLink on the GPU box, gloo in
the CPU tests); nothing in here touches the kernels.

Loss scaling: a loss that is a *mean* over the local batch must be rescaled so that the summed gradient equals the
single-process gradient of the mean over the global batch -- ``global_mean_scale`` gives that factor.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def world_info(group: Optional[dist.ProcessGroup] = None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``num_items`` for ``rank``; sizes differ by at most one, earlier ranks larger."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(num_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rays(origins: Tensor, directions: Tensor, rank: int, world_size: int, batch: Optional[int] = None):
    """This rank's rays.  ``batch=None``: one contiguous flat-index range (neighbouring rays stay together, which is what
    the kernels' coalescing wants).  ``batch=B``: whole B-ray batches dealt round-robin (batch b goes to rank b % world)."""
    assert origins.shape == directions.shape and origins.dim() == 2
    n = origins.shape[0]
    if batch is None:
        lo, hi = shard_bounds(n, rank, world_size)
        return origins[lo:hi], directions[lo:hi]
    starts = list(range(0, n, batch))[rank::world_size]
    if not starts:
        return origins[:0], directions[:0]
    idx = torch.cat([torch.arange(s, min(s + batch, n), device=origins.device) for s in starts])
    return origins[idx], directions[idx]


def shard_views(num_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin view indices for ``rank`` (cfg 4 of BASELINE.json: 100 views over 8 ranks -> 12 or 13 each)."""
    return list(range(rank, num_views, world_size))


def global_mean_scale(local_count: int, group: Optional[dist.ProcessGroup] = None) -> float:
    """Factor that turns a local-mean loss into this rank's share of the global-mean loss: local_count / global_count."""
    rank, world = world_info(group)
    if world == 1:
        return 1.0
    t = torch.tensor([float(local_count)], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    total = float(t.item())
    return float(local_count) / total if total > 0 else 0.0


class VoxelGradAllReducer:
    """Sums the voxel gradients of a replicated grid across ranks with a single collective.

    The gradients of all given parameters are gathered into one persistent flat buffer, reduced with one
    ``all_reduce(SUM)`` and scattered back -- one message of (F+1)*X*Y*Z floats per optimiser step instead of one per
    parameter.  ``reduce_flat`` is the same on a caller-owned buffer (e.g. the packed gradient volume of the C ABI).
    """

    def __init__(self, params: Iterable[Tensor], group: Optional[dist.ProcessGroup] = None) -> None:
        self.params: List[Tensor] = [p for p in params]
        self.group = group
        self._flat: Optional[Tensor] = None
        self.num_collectives = 0

    def _buffer(self) -> Tensor:
        numel = sum(p.numel() for p in self.params)
        ref = self.params[0]
        if self._flat is None or self._flat.numel() != numel or self._flat.device != ref.device:
            self._flat = torch.empty(numel, dtype=torch.float32, device=ref.device)
        return self._flat

    def reduce_flat(self, flat: Tensor) -> Tensor:
        rank, world = world_info(self.group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            self.num_collectives += 1
        return flat

    def __call__(self) -> None:
        """In-place: every ``p.grad`` becomes the sum over ranks (parameters without a gradient contribute zeros)."""
        rank, world = world_info(self.group)
        if world == 1 or not self.params:
            return
        flat = self._buffer()
        offset = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                flat[offset : offset + n].zero_()
            else:
                flat[offset : offset + n].copy_(p.grad.reshape(-1))
            offset += n
        self.reduce_flat(flat)
        offset = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = flat[offset : offset + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(flat[offset : offset + n].reshape(p.shape))
            offset += n
