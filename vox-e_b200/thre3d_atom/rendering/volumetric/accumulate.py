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


def _fused_only(name: str):
    def stage(*args, **kwargs):
        raise NotImplementedError(
            f"{name} is a stage of the reference's unfused render pipeline; in this package the sampler, point processor and "
            "accumulator are one CUDA kernel pair entered through thre3d_atom.thre3d_reprs.renderers.render_sh_voxel_grid"
            "[_attn] (there is deliberately no PyTorch fallback)"
        )

    stage.__name__ = stage.__qualname__ = name
    stage.__doc__ = f"Name kept for import compatibility (accumulate.py upstream); calling it raises NotImplementedError."
    return stage


accumulate_radiance_density_on_rays = _fused_only("accumulate_radiance_density_on_rays")  # accumulate.py:31-113 upstream
accumulate_radiance_density_on_rays_attn = _fused_only("accumulate_radiance_density_on_rays_attn")  # accumulate.py:115-198
