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

    def __init__(self, params: Iterable[Tensor], group: Optional[dist.ProcessGroup] = None, grids: Sequence = ()) -> None:
        self.params: List[Tensor] = [p for p in params]
        self.group = group
        self.grids = list(grids)  # VoxelGrids whose render gradients may be deferred (accumulate_render_gradients / FusedVoxelAdam)
        self._flat: Optional[Tensor] = None
        self.num_collectives = 0

    def _deferred_grids(self) -> list:
        """Grids that own one of ``self.params`` and keep render gradients in a packed sink volume: the ones handed to the
        constructor plus any other grid registered with ``voxe_b200.optim`` (so that a forgotten ``grids=`` cannot make the
        ranks step on local, unreduced render gradients)."""
        from voxe_b200 import optim

        owned = {id(p) for p in self.params}
        grids = list(self.grids)
        for grid in list(optim._tracked):
            if grid not in grids and (id(grid.densities) in owned or id(grid.features) in owned):
                grids.append(grid)
        return [g for g in grids if g.render_gradient_accumulator is not None]

    def reduce_deferred(self) -> int:
        """All-reduce the packed sink volumes of the deferred grids in place (ONE collective per grid on the volume the
        backward kernels scattered into; ``FusedVoxelAdam`` or the optimiser pre-step hook then consume the summed
        gradients).  Ranks whose volume is clean still take part (their zeros are part of the sum).  Returns the number
        of volumes reduced."""
        rank, world = world_info(self.group)
        n = 0
        for grid in self._deferred_grids():
            acc = grid.render_gradient_accumulator
            if acc.buffer is None:  # nothing rendered yet on this rank: allocate the (zero) volume so the collective matches
                spec = grid.fused_spec()
                acc.get(grid.packed_cache().get(spec, grid.densities, grid.features))
            if world > 1:
                peer = getattr(acc, "peer_volume", None)
                if peer is not None and acc.buffer is peer.buffer:  # the library's own kernel, in place on the sink volume
                    if acc.sparse_sink:
                        peer.allreduce_sparse(int(acc.touch_tag[0]))
                    else:
                        peer.allreduce()
                    self.num_collectives += 1
                else:
                    self.reduce_flat(acc.buffer)
                acc.dirty = True  # the sum may be non-zero even where this rank's share was
            n += 1
        return n

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
        """In-place: every ``p.grad`` becomes the sum over ranks (parameters without a gradient contribute zeros).
        Render gradients that are still deferred in a grid's packed sink volume are materialised into ``.grad`` first, so
        the one collective below carries them (use ``reduce_deferred`` to reduce the packed volumes themselves)."""
        rank, world = world_info(self.group)
        if world == 1 or not self.params:
            return
        for grid in self._deferred_grids():
            grid.materialize_render_gradients()
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


# Share of a launch's CTAs (out of 8) that take the multicast path when a multicast mapping exists (see csrc/voxe_collective.cu).
DEFAULT_MULTICAST_SHARE = 8


class PeerGradVolume:
    """A packed gradient volume that every rank of the group maps into its address space, reduced in place by the
    library's own kernel (``voxe_allreduce_grads_peer``: two-shot over NVLink peer memory, through the NVSwitch's
    multicast reduction when the mapping has one) instead of an NCCL call.

    The memory comes from torch's symmetric-memory allocator -- plumbing only: it allocates, exchanges the handles and
    maps the peers; the exchange itself is ``csrc/voxe_collective.cu``.  ``buffer`` is an ordinary fp32 CUDA tensor; hand
    it to the backward kernels as their gradient volume (``adopt``) and call ``allreduce()`` once per optimiser step."""

    def __init__(self, n_floats: int, device: torch.device, group: Optional[dist.ProcessGroup] = None, multicast: bool = True,
                 multicast_share: Optional[int] = None) -> None:
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from voxe_b200 import _native as nat

        group = group if group is not None else dist.group.WORLD
        self._nat, self._lib = nat, nat.load_library()
        self.group = group
        self.n_floats = int(n_floats)
        if self.n_floats % 4:
            raise ValueError("the volume must be a whole number of 16-byte vectors")
        self.buffer = symm_mem.empty(self.n_floats, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self._signals = symm_mem.empty(nat.SIGNAL_WORDS, dtype=torch.int32, device=device)
        self._signals.zero_()
        self._fail = torch.zeros(1, dtype=torch.int32, device=device)
        h_buf = symm_mem.rendezvous(self.buffer, group)
        h_sig = symm_mem.rendezvous(self._signals, group)
        self.rank, self.world_size = int(h_buf.rank), int(h_buf.world_size)
        if self.world_size > nat.MAX_PEERS:
            raise ValueError(f"at most {nat.MAX_PEERS} ranks")
        desc = nat.VoxePeerDesc()
        desc.world_size, desc.rank = self.world_size, self.rank
        for k in range(self.world_size):
            desc.buffers[k] = int(h_buf.buffer_ptrs[k])
            desc.signals[k] = int(h_sig.buffer_ptrs[k])
        mc = int(h_buf.multicast_ptr) if multicast else 0
        desc.multicast = mc if mc else None
        self.multicast = bool(mc)
        # of every 8 CTAs, how many go through the switch's multicast reduction (the rest: plain peer loads / stores)
        self.multicast_share = (DEFAULT_MULTICAST_SHARE if multicast_share is None else int(multicast_share)) if mc else 0
        desc.multicast_share = self.multicast_share
        self._desc, self._handles = desc, (h_buf, h_sig)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's zero-fill of its signal pad is done before anybody's first launch

    def allreduce(self) -> None:
        """buffer <- sum over ranks, in place, on the current stream (one launch; every rank must call it)."""
        dev = self.buffer.device
        with torch.cuda.device(dev):
            self._nat.check(self._lib.voxe_allreduce_grads_peer(self._desc, self.n_floats, self._fail.data_ptr(),
                                                                torch.cuda.current_stream(dev).cuda_stream), "voxe_allreduce_grads_peer")

    def enable_sparse(self, gspec) -> Tensor:
        """Allocate (collectively, on every rank) the peer-mapped brick-flag array of ``gspec``'s packed volume and return it:
        pass it to the backward kernels as ``touched`` (``voxe_render_bwd``), then :meth:`allreduce_sparse` exchanges only the
        bricks some rank touched.  The volume must be the packed gradient volume of that grid."""
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        gd = gspec.to_native()
        if int(self._lib.voxe_packed_floats(gd)) != self.n_floats:
            raise ValueError("the peer volume is not the packed gradient volume of this grid")
        n = int(self._lib.voxe_peer_touched_bytes(gd))
        self.touched = symm_mem.empty(n, dtype=torch.uint8, device=self.buffer.device)
        self.touched.zero_()
        h = symm_mem.rendezvous(self.touched, self.group)
        self._touched_ptrs = (ctypes.c_void_p * self.world_size)(*[int(h.buffer_ptrs[k]) for k in range(self.world_size)])
        self._handles = self._handles + (h,)
        self._gd = gd
        torch.cuda.synchronize(self.buffer.device)
        dist.barrier(self.group)
        return self.touched

    def allreduce_sparse(self, tag: int) -> None:
        """buffer <- sum over ranks on the bricks whose flag carries ``tag`` on some rank (``voxe_allreduce_grads_peer_sparse``;
        two launches on the current stream, every rank with the same tag); afterwards ``touched`` holds the union."""
        dev = self.buffer.device
        with torch.cuda.device(dev):
            self._nat.check(self._lib.voxe_allreduce_grads_peer_sparse(self._desc, self._gd, self._touched_ptrs, int(tag), self._fail.data_ptr(),
                                                                       torch.cuda.current_stream(dev).cuda_stream), "voxe_allreduce_grads_peer_sparse")

    def failed(self) -> bool:
        """True when a launch gave up waiting for a peer (synchronises)."""
        return bool(self._fail.item())

    def adopt(self, accumulator, sparse_spec=None) -> None:
        """Make this volume the gradient volume of a ``PackedGradAccumulator`` (a grid's deferred-gradient sink).
        ``sparse_spec`` (the grid's ``fused_spec()``): also keep the brick-flag trail in peer-mapped memory, so that
        ``VoxelGradAllReducer.reduce_deferred()`` exchanges only the bricks some rank touched in the step
        (``allreduce_sparse``) and the hand-over into ``.grad`` follows the union -- for grids of which a step writes a small
        part.  Collective when ``sparse_spec`` is given."""
        accumulator.buffer = self.buffer
        accumulator.touched = None
        accumulator.sparse_sink = False
        accumulator.dirty = False
        accumulator.peer_volume = self
        if sparse_spec is not None:
            touched = getattr(self, "touched", None)
            if touched is None:
                touched = self.enable_sparse(sparse_spec)
            accumulator.enable_trail(sparse_spec, self.buffer, touched=touched)


class PeerGradients:
    """Dense ``.grad`` tensors of a replicated grid, laid out back to back in ONE peer-mapped buffer, so that the step's
    collective is a single in-place launch of the library's all-reduce kernel on the very memory autograd accumulates into --
    no flat staging copy, no second collective for the second parameter.

        grads = PeerGradients([grid.densities, grid.features])      # once, collectively, on every rank
        for step in ...:
            grads.zero()                                             # instead of optimizer.zero_grad(): keeps the views
            loss.backward()                                          # the render's node adds into p.grad (these views)
            grads.allreduce()                                        # ONE collective: p.grad = sum over ranks
            optimizer.step()

    ``optimizer.zero_grad(set_to_none=False)`` keeps the views as well; ``set_to_none=True`` (torch's default) drops them --
    call :meth:`attach` again afterwards, or use :meth:`zero`."""

    def __init__(self, params: Sequence[Tensor], group: Optional[dist.ProcessGroup] = None, multicast: Optional[bool] = None) -> None:
        self.params = list(params)
        numel = sum(p.numel() for p in self.params)
        padded = (numel + 3) // 4 * 4  # whole 16-byte vectors
        rank, world = world_info(group)
        use_multicast = (world > 4) if multicast is None else multicast  # measured crossover: see PeerGradVolume / DESIGN.md section 7
        self.volume = PeerGradVolume(padded, self.params[0].device, group=group, multicast=use_multicast)
        self.views: List[Tensor] = []
        offset = 0
        for p in self.params:
            self.views.append(self.volume.buffer[offset : offset + p.numel()].view(p.shape))
            offset += p.numel()
        self.attach()

    def attach(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v

    def zero(self) -> None:
        self.attach()
        self.volume.buffer.zero_()

    def allreduce(self) -> None:
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                raise RuntimeError("a parameter's .grad is no longer the peer-mapped view (zero_grad(set_to_none=True)?); call attach() / zero()")
        self.volume.allreduce()
