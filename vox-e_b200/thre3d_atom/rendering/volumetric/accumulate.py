"""Occupancy model named by render configurations.

``density2occupancy_pb`` is pickled *by reference* inside every checkpoint's ``render_config`` (volumetric_model.py:95),
so it must stay importable from this path.  In this package it doubles as a marker: the fused kernels implement exactly
this law (alpha = 1 - exp(-sigma * delta), accumulate.py:24-28 upstream) and ``render_sh_voxel_grid`` refuses any other
callable instead of silently running something else.
"""
import torch
from torch import Tensor


def density2occupancy_pb(densities: Tensor, deltas: Tensor) -> Tensor:
    """Beer-Lambert occupancy of a segment of length ``deltas`` with density ``densities``; in [0, 1) for density >= 0."""
    return 1.0 - torch.exp(-(densities * deltas))
