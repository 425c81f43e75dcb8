"""Value types of the render interface (reference: thre3d_atom/rendering/volumetric/render_interface.py:13-131).

``Rays`` / ``RenderOut`` / ``RenderOutAttn`` keep the reference's field names, shape checks and helpers.  The three-stage
``render()`` driver of the reference (:140-171) has no counterpart here: sampler, point processor and accumulator are
one fused CUDA kernel (``voxe_b200.fused_render``), entered from ``thre3d_atom.thre3d_reprs.renderers``.
"""
import dataclasses
from typing import Any, Callable, Dict, NamedTuple, Optional

import torch
from torch import Tensor

from thre3d_atom.utils.constants import NUM_ATTN_CHANNELS, NUM_COLOUR_CHANNELS, NUM_COORD_DIMENSIONS
from thre3d_atom.utils.imaging_utils import CameraBounds

ExtraInfo = Dict[str, Any]


@dataclasses.dataclass
class Rays:
    origins: Tensor  # [..., 3]
    directions: Tensor  # [..., 3]   (not normalised)

    def __post_init__(self):
        assert self.origins.shape == self.directions.shape, "ray-origins and ray-directions are incompatible :("
        assert self.origins.shape[-1] == self.directions.shape[-1] == NUM_COORD_DIMENSIONS, (
            "Sorry, we only support 3D coordinate-spaces at the moment. Please cast your rays in 3 dimensions only :D"
        )

    def __getitem__(self, item) -> "Rays":
        return Rays(origins=self.origins[item, :], directions=self.directions[item, :])

    def __len__(self) -> int:
        return len(self.origins)

    def to(self, device: torch.device) -> "Rays":
        return Rays(self.origins.to(device), self.directions.to(device))


class _RenderedMaps:
    """detach()/to() shared by the two output types; ``_main`` names the primary map."""

    _main = "colour"

    def _rebuild(self, fn: Callable[[Tensor], Tensor]):
        return type(self)(
            **{self._main: fn(getattr(self, self._main))},
            depth=fn(self.depth),
            extra={key: fn(value) for key, value in self.extra.items()},
        )

    def detach(self):
        return self._rebuild(lambda t: t.detach())

    def to(self, device: torch.device):
        return self._rebuild(lambda t: t.to(device))

    @classmethod
    def _trusted(cls, main: Tensor, depth: Tensor, extra: ExtraInfo):
        """Build from maps whose shapes the fused kernels guarantee ([R, C], [R, 1]): skips the shape asserts of the public
        constructor (a few microseconds on a path whose whole host budget is ~40)."""
        out = object.__new__(cls)
        setattr(out, cls._main, main)
        out.depth = depth
        out.extra = extra
        return out

    def _check(self, channels: int) -> None:
        main = getattr(self, self._main)
        assert main.shape[:-1] == self.depth.shape[:-1], "rendered colour maps and depth maps are shape-incompatible"
        assert main.shape[-1] == channels, f"expected {channels} channel(s) in the rendered {self._main} map"
        assert self.depth.shape[-1] == 1, "Sorry, depth map should only have 1 dimensional data channel"
        if self.extra is None:
            self.extra = {}


@dataclasses.dataclass
class RenderOut(_RenderedMaps):
    colour: Tensor  # [..., 3]
    depth: Tensor  # [..., 1]
    extra: Optional[ExtraInfo] = None

    def __post_init__(self):
        self._check(NUM_COLOUR_CHANNELS)


@dataclasses.dataclass
class RenderOutAttn(_RenderedMaps):
    attn: Tensor  # [..., 1]
    depth: Tensor  # [..., 1]
    extra: Optional[ExtraInfo] = None
    _main = "attn"

    def __post_init__(self):
        self._check(NUM_ATTN_CHANNELS)


class SampledPointsOnRays(NamedTuple):
    points: Tensor  # [N, num_samples, 3]
    depths: Tensor  # [N, num_samples]


ProcessedPointsOnRays = SampledPointsOnRays

RaySamplerFunction = Callable[[Rays, CameraBounds, int], SampledPointsOnRays]
PointProcessorFunction = Callable[[SampledPointsOnRays, Rays], ProcessedPointsOnRays]
AccumulatorFunction = Callable[[ProcessedPointsOnRays, Rays], RenderOut]


def render(rays: Rays, camera_bounds: CameraBounds, num_samples: int, sampler_fn: RaySamplerFunction,
           point_processor_fn: PointProcessorFunction, accumulator_fn: AccumulatorFunction) -> RenderOut:
    """The reference's three-stage driver (render_interface.py:140-171 upstream).  Name kept for import compatibility: the
    stages are fused into one kernel pair here, so there is no pipeline of callables to drive."""
    raise NotImplementedError(
        "render() drives the reference's unfused sampler -> processor -> accumulator callables; this package renders through "
        "thre3d_atom.thre3d_reprs.renderers.render_sh_voxel_grid (one fused CUDA launch, no PyTorch fallback)"
    )


def render_attn(rays: Rays, camera_bounds: CameraBounds, num_samples: int, sampler_fn: RaySamplerFunction,
                point_processor_fn: PointProcessorFunction, accumulator_fn: AccumulatorFunction) -> RenderOutAttn:
    """Attention twin of ``render`` (render_interface.py:174-205 upstream); see there."""
    raise NotImplementedError(
        "render_attn() drives the reference's unfused callables; this package renders through "
        "thre3d_atom.thre3d_reprs.renderers.render_sh_voxel_grid_attn"
    )
