"""CPU oracle for the per-step grid regularisers of Vox-E's edit loop.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file; the product
(``vox-e_b200/voxe_b200/regularizers.py`` -> ``csrc/voxe_regularizers.cu``) never does.

Restated from the reference's formulas, not its code (neighbour slices instead of ``Tensor.diff``, the correlation written
through its normalised moments), so that a slip in either shows up as a disagreement:

  * tv_loss ................. ``_tv_loss_on_grid``           thre3d_atom/modules/sds_trainer.py:563-567
                              (same function: attn_grid_trainer.py:659-663, grid_refine.py:709-713), with the ReLU the
                              trainer applies to ``_densities`` first (sds_trainer.py:318-322)
  * density_pair_loss ....... ``density_correlation_loss_fn`` thre3d_atom/modules/sds_trainer.py:494-505 and
                              ``_density_correlation_loss``   thre3d_atom/modules/sds_trainer.py:507-524

Parity pin: ``tests/golden/make_golden_regularizers.py`` executes the reference's own function bodies (extracted from
``/root/reference/thre3d_atom/modules/sds_trainer.py`` with ``ast``; the module itself imports diffusers / wandb, which
are not installed) and writes ``tests/golden/regularizers.npz``; ``tests/test_regularizers.py`` checks this file against
it.  Gradients come from torch autograd over the restatement; ``dtype=torch.float64`` is the truth mode.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

CORRELATION_EPS = 0.0000001  # sds_trainer.py:509


def tv_loss(grid: Tensor, relu: bool = False) -> Tensor:
    """Mean absolute forward difference along x, y and z of a channel-last grid [X,Y,Z,C], averaged over the axes."""
    h = grid * (grid > 0) if relu else grid  # ReLU with derivative 0 AT 0, like torch.nn.ReLU's backward (result > 0)
    along_x = (h[1:, :, :, :] - h[:-1, :, :, :]).abs()
    along_y = (h[:, 1:, :, :] - h[:, :-1, :, :]).abs()
    along_z = (h[:, :, 1:, :] - h[:, :, :-1, :]).abs()
    return (along_x.sum() / along_x.numel() + along_y.sum() / along_y.numel() + along_z.sum() / along_z.numel()) / 3.0


def density_pair_loss(a: Tensor, b: Tensor, mode: str = "correlation") -> Tensor:
    """Loss between the edited density grid ``a`` and the frozen pretrained grid ``b``."""
    n = a.numel()
    if mode == "l2":
        return ((a - b) ** 2).sum() / n
    if mode == "l1":
        return (a - b).abs().sum() / n
    da, db = a - a.sum() / n, b - b.sum() / n
    var_a, var_b = (da * da).sum() / n, (db * db).sum() / n
    cov = (da * db).sum() / n
    return 1.0 - cov / (torch.sqrt(var_a * var_b) + CORRELATION_EPS)


def correlation_grid(a: Tensor, b: Tensor) -> Tensor:
    """Second return value of ``_density_correlation_loss``: the per-voxel covariance over (denominator + eps)."""
    n = a.numel()
    da, db = a - a.sum() / n, b - b.sum() / n
    denom = torch.sqrt((da * da).sum() / n * ((db * db).sum() / n))
    return da * db / (denom + CORRELATION_EPS)


def with_grad(fn, x: Tensor, *args, upstream: float = 1.0, dtype=torch.float64, **kwargs) -> Dict[str, Tensor]:
    """loss and upstream * dloss/dx of ``fn(x, *args, **kwargs)`` in ``dtype``."""
    xx = x.detach().to(dtype).clone().requires_grad_(True)
    rest = [t.detach().to(dtype) if isinstance(t, Tensor) else t for t in args]
    loss = fn(xx, *rest, **kwargs)
    (g,) = torch.autograd.grad(loss * upstream, xx)
    return {"loss": loss.detach(), "grad": g}
