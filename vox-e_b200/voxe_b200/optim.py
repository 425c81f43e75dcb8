"""The optimiser-step side of the render path.

Once a render takes tens of microseconds, the full-grid passes around it bound a training iteration: per call the
reference-style path zero-fills a packed gradient volume, unpacks it, and autograd adds it into ``.grad``; per step
``torch.optim.Adam`` reads and writes parameters and moments in ~10 kernels, and the next render repacks the grid.

* ``VoxelGrid.accumulate_render_gradients()`` defers gradient materialisation: backward kernels scatter into one
  persistent packed volume.  A global optimiser pre-step hook (registered here) materialises it into ``.grad`` right
  before *any* ``torch.optim`` optimiser steps, so stock ``torch.optim.Adam`` loops such as the reference's
  (modules/trainers.py:247-255,348-351) keep working unchanged.
* ``FusedVoxelAdam`` goes further: one kernel (``voxe_adam_step``) consumes the packed gradients (+ any dense ``.grad``
  from torch-side losses like the TV / density-correlation terms of modules/sds_trainer.py:290-326), applies the Adam
  update to parameters and moments, refreshes the packed volume for the next render and zeroes the gradient volume.
  It is a ``torch.optim.Optimizer`` (param groups, ``state_dict`` keys ``step / exp_avg / exp_avg_sq`` like
  ``torch.optim.Adam``), so ``ExponentialLR`` schedulers and checkpointing code apply unchanged.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch
from torch.optim import Optimizer

from voxe_b200 import _native as nat
from voxe_b200.render_function import _stream_ptr

_tracked = weakref.WeakSet()
_hook_handle = None


def track_grid(grid) -> None:
    """Remember a grid with deferred gradients; installs the global pre-step hook on first use."""
    global _hook_handle
    _tracked.add(grid)
    if _hook_handle is None:
        from torch.optim.optimizer import register_optimizer_step_pre_hook

        _hook_handle = register_optimizer_step_pre_hook(_materialize_before_step)


def _materialize_before_step(optimizer, args, kwargs) -> None:
    if isinstance(optimizer, FusedVoxelAdam):
        return  # consumes the packed gradients itself
    owned = {id(p) for group in optimizer.param_groups for p in group["params"]}
    for grid in list(_tracked):
        if id(grid.densities) in owned or id(grid.features) in owned:
            grid.materialize_render_gradients()


class FusedVoxelAdam(Optimizer):
    """Adam on a ``VoxelGrid``'s ``_densities`` and ``_features`` in one fused CUDA pass per step."""

    def __init__(self, voxel_grid, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"Invalid betas: {betas}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        params = [voxel_grid.densities, voxel_grid.features]
        if not all(isinstance(p, torch.nn.Parameter) for p in params):
            raise ValueError("FusedVoxelAdam needs a tunable VoxelGrid (densities / features must be Parameters)")
        super().__init__([{"params": params}], dict(lr=lr, betas=betas, eps=eps))
        self._grid = weakref.ref(voxel_grid)
        voxel_grid.accumulate_render_gradients(True)

    def zero_grad(self, set_to_none: bool = True) -> None:
        super().zero_grad(set_to_none=set_to_none)
        grid = self._grid()
        if grid is not None and grid.render_gradient_accumulator is not None:
            grid.render_gradient_accumulator.zero()  # no-op when step() already consumed it

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grid = self._grid()
        if grid is None:
            raise RuntimeError("the VoxelGrid of this optimiser no longer exists")
        group = self.param_groups[0]
        dens, feat = group["params"]
        if dens is not grid.densities or feat is not grid.features:
            raise RuntimeError("the grid's parameters were replaced after the optimiser was built; build a new optimiser")
        if dens.device.type != "cuda":
            raise RuntimeError("FusedVoxelAdam runs on CUDA only (no CPU fallback)")
        moments = []
        for p in (dens, feat):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] += 1
            moments.append(st)
        step = int(moments[0]["step"].item())
        spec = grid.fused_spec()
        gd = spec.to_native()
        acc = grid.render_gradient_accumulator
        packed_grad = acc.buffer if (acc is not None and acc.dirty) else None
        cache = grid.packed_cache()
        packed = cache.peek()
        adam = nat.VoxeAdamDesc(lr=float(group["lr"]), beta1=float(group["betas"][0]), beta2=float(group["betas"][1]),
                                eps=float(group["eps"]), step=step)
        dense = [None if p.grad is None else p.grad.contiguous() for p in (dens, feat)]
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        lib = nat.load_library()
        with torch.cuda.device(dens.device):
            nat.check(
                lib.voxe_adam_step(gd, adam, dens.data_ptr(), feat.data_ptr(), ptr(packed), ptr(packed_grad), ptr(dense[0]), ptr(dense[1]),
                                   moments[0]["exp_avg"].data_ptr(), moments[0]["exp_avg_sq"].data_ptr(),
                                   moments[1]["exp_avg"].data_ptr(), moments[1]["exp_avg_sq"].data_ptr(), _stream_ptr(dens.device)),
                "voxe_adam_step",
            )
        if acc is not None:
            acc.dirty = False  # the kernel zeroed what it consumed
        # the parameters changed behind autograd's back: bump their version counters, then vouch for the packed copy
        torch.autograd.graph.increment_version([dens, feat])
        cache.mark_fresh(spec, dens, feat)
        return loss
