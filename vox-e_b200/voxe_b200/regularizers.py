"""Per-step full-grid regularisers of Vox-E's edit loop on the streaming kernels of ``csrc/voxe_regularizers.cu``
(SURVEY.md row f2).  Same names, arguments and return values as the functions they replace:

    density_correlation_loss_fn   thre3d_atom/modules/sds_trainer.py:494-505   (weight 200 by default in the edit script)
    _density_correlation_loss     thre3d_atom/modules/sds_trainer.py:507-524
    _tv_loss_on_grid              thre3d_atom/modules/sds_trainer.py:563-567, attn_grid_trainer.py:659-663,
                                  grid_refine.py:709-713

so a trainer switches over with ``from voxe_b200.regularizers import density_correlation_loss_fn, _tv_loss_on_grid``
(INTEGRATION.md).  Each loss is a differentiable 0-dim tensor: the forward is one read of the grid(s) and a small
reduction, the backward one more read that writes the dense gradient autograd expects.  The ``accumulate_*`` variants
skip autograd altogether for callers that only need "add weight * dloss/dgrid into .grad and tell me the loss": one pass
for the TV term, two for the correlation term, nothing allocated -- the dense gradients they leave in ``.grad`` are what
``FusedVoxelAdam`` / ``voxe_adam_step`` consume beside the render's packed gradient volume.

CUDA only, fp32 only, no fallback: CPU tensors raise, as everywhere in this package.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from voxe_b200 import _native as nat
from voxe_b200.render_function import _require_cuda, _stream_ptr

_workspaces = {}


def _workspace(dev: torch.device) -> Tensor:
    """One small reduction workspace per (device, stream): calls on one stream are ordered, so they can share it."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = torch.zeros(nat.REG_WORKSPACE_DOUBLES, dtype=torch.float64, device=dev)
    return ws


def _grid_args(grid: Tensor) -> Tuple[Tensor, "ctypes.Array", int]:
    if grid.dim() != 4:
        raise ValueError(f"expected a channel-last grid [X, Y, Z, C] (got shape {tuple(grid.shape)})")
    _require_cuda(grid)
    g = grid.detach().contiguous()
    dims = (ctypes.c_int32 * 3)(*g.shape[:3])
    return g, dims, int(g.shape[3])


def _tv_call(g: Tensor, dims, channels: int, relu: bool, loss: Optional[Tensor], upstream: Optional[Tensor], scale: float,
             grad: Optional[Tensor], accumulate: bool) -> None:
    dev = g.device
    lib = nat.load_library()
    with torch.cuda.device(dev):
        ws = _workspace(dev) if loss is not None else None
        nat.check(
            lib.voxe_tv_regularizer(g.data_ptr(), ctypes.byref(dims), channels, int(relu), None if ws is None else ws.data_ptr(),
                                    None if loss is None else loss.data_ptr(), None if upstream is None else upstream.data_ptr(),
                                    float(scale), None if grad is None else grad.data_ptr(), int(accumulate), _stream_ptr(dev)),
            "voxe_tv_regularizer",
        )


class _TVLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid: Tensor, relu: bool) -> Tensor:
        g, dims, channels = _grid_args(grid)
        loss = torch.empty((), dtype=torch.float32, device=g.device)
        _tv_call(g, dims, channels, relu, loss, None, 1.0, None, False)
        ctx.save_for_backward(grid)
        ctx.relu = relu
        return loss

    @staticmethod
    def backward(ctx, g_loss: Tensor):
        (grid,) = ctx.saved_tensors
        g, dims, channels = _grid_args(grid)
        grad = torch.empty_like(g)
        up = g_loss.detach().to(device=g.device, dtype=torch.float32).contiguous()
        _tv_call(g, dims, channels, ctx.relu, None, up, 1.0, grad, False)
        return grad.view_as(grid), None


def _tv_loss_on_grid(grid: Tensor, relu: bool = False) -> Tensor:
    """``(grid.diff(0).abs().mean() + grid.diff(1).abs().mean() + grid.diff(2).abs().mean()) / 3`` of a channel-last grid
    [X,Y,Z,C] (sds_trainer.py:563-567).  ``relu=True`` folds in the ``torch.nn.ReLU()`` the trainer applies to
    ``_densities`` first (sds_trainer.py:319-321), saving the activated copy of the grid."""
    return _TVLoss.apply(grid, bool(relu))


def accumulate_tv_gradient(param: Tensor, weight: float, relu: bool = False) -> Tensor:
    """``param.grad += weight * d tv_loss(param) / d param`` and return the (detached, unweighted) loss -- loss and gradient
    in ONE pass over the grid, no autograd graph.  Equivalent to ``(weight * _tv_loss_on_grid(param)).backward()``."""
    g, dims, channels = _grid_args(param)
    if g.data_ptr() != param.data_ptr():
        raise ValueError("accumulate_tv_gradient needs a contiguous parameter")
    if param.grad is None:
        param.grad = torch.zeros_like(param, memory_format=torch.contiguous_format)
    loss = torch.empty((), dtype=torch.float32, device=g.device)
    _tv_call(g, dims, channels, relu, loss, None, float(weight), _dense_grad(param), True)
    return loss


def _dense_grad(param: Tensor) -> Tensor:
    grad = param.grad
    if grad.dtype != torch.float32 or not grad.is_contiguous() or grad.shape != param.shape or grad.device != param.device:
        raise ValueError("the parameter's .grad must be a dense contiguous fp32 tensor shaped like the parameter")
    return grad


_MODES = {"correlation": nat.PAIR_CORRELATION, "l2": nat.PAIR_L2, "l1": nat.PAIR_L1}


def _pair_args(a: Tensor, b: Tensor) -> Tuple[Tensor, Tensor]:
    if a.shape != b.shape:
        raise ValueError(f"the two density grids must have one shape (got {tuple(a.shape)} and {tuple(b.shape)})")
    _require_cuda(a, b)
    return a.detach().contiguous(), b.detach().contiguous()


def _pair_forward(a: Tensor, b: Tensor, mode: int, want_grid: bool, ws: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
    dev = a.device
    lib = nat.load_library()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    corr = torch.empty_like(a) if (want_grid and mode == nat.PAIR_CORRELATION) else None
    with torch.cuda.device(dev):
        nat.check(lib.voxe_pair_loss(a.data_ptr(), b.data_ptr(), a.numel(), mode, ws.data_ptr(), loss.data_ptr(),
                                     None if corr is None else corr.data_ptr(), _stream_ptr(dev)), "voxe_pair_loss")
    return loss, corr


def _pair_backward(a: Tensor, b: Tensor, mode: int, ws: Tensor, upstream: Optional[Tensor], scale: float, grad: Tensor,
                   accumulate: bool) -> None:
    dev = a.device
    lib = nat.load_library()
    with torch.cuda.device(dev):
        nat.check(lib.voxe_pair_loss_grad(a.data_ptr(), b.data_ptr(), a.numel(), mode, ws.data_ptr(),
                                          None if upstream is None else upstream.data_ptr(), float(scale), grad.data_ptr(),
                                          int(accumulate), _stream_ptr(dev)), "voxe_pair_loss_grad")


class _PairLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sds_density: Tensor, regular_density: Tensor, mode: int, want_grid: bool):
        a, b = _pair_args(sds_density, regular_density)
        ws = torch.zeros(nat.REG_WORKSPACE_DOUBLES, dtype=torch.float64, device=a.device)  # its statistics are kept for the backward
        loss, corr = _pair_forward(a, b, mode, want_grid, ws)
        ctx.save_for_backward(sds_density, regular_density, ws)
        ctx.mode = mode
        if corr is None:
            return loss
        corr = corr.view_as(sds_density)
        ctx.mark_non_differentiable(corr)
        return loss, corr

    @staticmethod
    def backward(ctx, g_loss: Tensor, *unused):
        sds_density, regular_density, ws = ctx.saved_tensors
        a, b = _pair_args(sds_density, regular_density)
        grad = torch.empty_like(a)
        up = g_loss.detach().to(device=a.device, dtype=torch.float32).contiguous()
        _pair_backward(a, b, ctx.mode, ws, up, 1.0, grad, False)
        return grad.view_as(sds_density), None, None, None


def _density_correlation_loss(sds_density: Tensor, regular_density: Tensor, return_correlation_grid: bool = True):
    """``1 - mean(correlation_grid)`` and the detached ``correlation_grid`` (sds_trainer.py:507-524).  The gradient flows
    to ``sds_density`` only: ``regular_density`` is the frozen pretrained grid (sds_trainer.py:139).  Nothing downstream
    reads the grid (``_feature_correlation_loss`` ignores its ``density_cov_grid`` argument, sds_trainer.py:526-534):
    pass ``return_correlation_grid=False`` to skip writing it and get ``(loss, None)``."""
    if return_correlation_grid:
        loss, corr = _PairLoss.apply(sds_density, regular_density, nat.PAIR_CORRELATION, True)
        return loss, corr
    return _PairLoss.apply(sds_density, regular_density, nat.PAIR_CORRELATION, False), None


def density_correlation_loss_fn(sds_density: Tensor, regular_density: Tensor, l2_mode: bool = False, l1_mode: bool = False,
                                return_correlation_grid: bool = True):
    """sds_trainer.py:494-505: ``mse_loss`` / ``l1_loss`` between the grids in the L2 / L1 modes (second value None),
    the correlation loss otherwise."""
    if l2_mode:
        return _PairLoss.apply(sds_density, regular_density, nat.PAIR_L2, False), None
    if l1_mode:
        return _PairLoss.apply(sds_density, regular_density, nat.PAIR_L1, False), None
    return _density_correlation_loss(sds_density, regular_density, return_correlation_grid)


def accumulate_density_loss_gradient(param: Tensor, regular_density: Tensor, weight: float, mode: str = "correlation") -> Tensor:
    """``param.grad += weight * dloss/dparam`` for the density loss of ``mode`` ("correlation" | "l2" | "l1") and return the
    detached, unweighted loss: two streaming passes (statistics, gradient), no autograd graph, nothing allocated but the
    0-dim loss.  Equivalent to ``(weight * density_correlation_loss_fn(param, regular_density, ...)[0]).backward()``."""
    a, b = _pair_args(param, regular_density)
    if a.data_ptr() != param.data_ptr():
        raise ValueError("accumulate_density_loss_gradient needs a contiguous parameter")
    if param.grad is None:
        param.grad = torch.zeros_like(param, memory_format=torch.contiguous_format)
    ws = _workspace(a.device)
    loss, _ = _pair_forward(a, b, _MODES[mode], False, ws)
    _pair_backward(a, b, _MODES[mode], ws, None, float(weight), _dense_grad(param), True)
    return loss
