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
        # the extra keys are torch.optim.Adam's remaining hyper-parameters at the only values this kernel implements, so
        # that state dicts move freely between the two optimisers
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                        differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__([{"params": params}], defaults)
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
        spec = grid.fused_spec()
        gd = spec.to_native()
        cache = grid.packed_cache()
        packed = cache.get(spec, dens, feat)  # no-op when the volume already mirrors the parameters
        acc = grid.render_gradient_accumulator
        packed_grad = acc.buffer if (acc is not None and acc.dirty) else None
        dense = [None if (p.grad is None or not p.requires_grad) else p.grad.contiguous() for p in (dens, feat)]
        if packed_grad is None and dense[0] is None and dense[1] is None:
            return loss  # no gradient anywhere: like torch.optim.Adam, which skips parameters whose .grad is None
        # a frozen tensor (requires_grad == False) is passed as NULL: the kernel leaves it and its moments alone
        live = [p if p.requires_grad else None for p in (dens, feat)]
        if live[0] is None and live[1] is None:
            if acc is not None:
                acc.zero()
            return loss
        st = self.state[dens]  # one shared record: the moments live in the packed layout, next to the packed volume
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["packed_exp_avg"] = torch.zeros_like(packed)
            st["packed_exp_avg_sq"] = torch.zeros_like(packed)
        st["step"] += 1
        step = int(st["step"].item())
        adam = nat.VoxeAdamDesc(lr=float(group["lr"]), beta1=float(group["betas"][0]), beta2=float(group["betas"][1]),
                                eps=float(group["eps"]), step=step)
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        lib = nat.load_library()
        with torch.cuda.device(dens.device):
            nat.check(
                lib.voxe_adam_step(gd, adam, ptr(live[0]), ptr(live[1]), packed.data_ptr(), ptr(packed_grad), ptr(dense[0]),
                                   ptr(dense[1]), st["packed_exp_avg"].data_ptr(), st["packed_exp_avg_sq"].data_ptr(),
                                   _stream_ptr(dens.device)),
                "voxe_adam_step",
            )
        if acc is not None:
            acc.dirty = False  # the kernel zeroed what it consumed
            if acc.sparse_sink:
                acc._next_tag()  # the next step leaves a fresh trail
        # the parameters changed behind autograd's back: bump their version counters, then vouch for the packed copy
        torch.autograd.graph.increment_version([dens, feat])
        cache.mark_fresh(spec, dens, feat)
        return loss

    # -- checkpoints keep torch.optim.Adam's per-parameter layout (exp_avg / exp_avg_sq shaped like the parameters) -------
    def moments(self):
        """(exp_avg, exp_avg_sq) for densities and features in the reference layout: ((m_d, v_d), (m_f, v_f))."""
        grid = self._grid()
        dens, feat = self.param_groups[0]["params"]
        st = self.state[dens]
        if len(st) == 0:
            z = lambda p: torch.zeros_like(p, memory_format=torch.contiguous_format)  # noqa: E731
            return (z(dens), z(dens)), (z(feat), z(feat))
        gd = grid.fused_spec().to_native()
        lib = nat.load_library()
        out = []
        with torch.cuda.device(dens.device):
            for key in ("packed_exp_avg", "packed_exp_avg_sq"):
                d, f = torch.empty_like(dens, memory_format=torch.contiguous_format), torch.empty_like(feat, memory_format=torch.contiguous_format)
                nat.check(lib.voxe_unpack_grad(gd, st[key].data_ptr(), d.data_ptr(), f.data_ptr(), 0, _stream_ptr(dens.device)), "voxe_unpack_grad")
                out.append((d, f))
        return (out[0][0], out[1][0]), (out[0][1], out[1][1])

    def state_dict(self):
        sd = super().state_dict()
        dens, feat = self.param_groups[0]["params"]
        if len(self.state[dens]):
            (m_d, v_d), (m_f, v_f) = self.moments()
            step = self.state[dens]["step"].clone()
            sd["state"] = {0: {"step": step, "exp_avg": m_d, "exp_avg_sq": v_d}, 1: {"step": step.clone(), "exp_avg": m_f, "exp_avg_sq": v_f}}
        return sd

    def load_state_dict(self, state_dict):
        from voxe_b200.render_function import pack_volume

        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        if 0 in state and 1 in state:
            grid = self._grid()
            dens, feat = self.param_groups[0]["params"]
            spec = grid.fused_spec()
            to = lambda t, like: t.to(device=like.device, dtype=torch.float32).reshape(like.shape).contiguous()  # noqa: E731
            self.state[dens] = {
                "step": torch.as_tensor(float(state[0]["step"]), dtype=torch.float32),
                "packed_exp_avg": pack_volume(spec, to(state[0]["exp_avg"], dens), to(state[1]["exp_avg"], feat)),
                "packed_exp_avg_sq": pack_volume(spec, to(state[0]["exp_avg_sq"], dens), to(state[1]["exp_avg_sq"], feat)),
            }
