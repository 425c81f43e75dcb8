"""VoxelGrid: the 3-D representation rendered by the fused kernels.

Keeps the reference's public surface (thre3d_atom/thre3d_reprs/voxels.py:19-517): constructor keywords, the
``_densities`` / ``_features`` / ``attn`` parameters and their state-dict keys, shape-checked property setters, the
AABB, config dictionaries for checkpoints, ``test_inside_volume`` and the rescale / load helpers.

What differs is where the arithmetic lives.  The reference's ``forward`` (voxels.py:287-342) runs two ``grid_sample``
calls plus a full-grid ``densities * scale`` pass per call; here the grid only *describes* itself to the kernels
(:meth:`fused_spec`, :meth:`packed_cache`) and the trilinear fetch happens inside ``voxe_render_fwd/bwd`` (or, for a stand-alone ``grid(points)`` query, inside
``voxe_query_points``).  Density
activations must therefore come from the closed set the kernels fuse -- pre in {Identity, abs}, post in {Identity,
ReLU, Softplus(beta=1, threshold=20)} -- and feature activations must be Identity (true for every script in the
reference); anything else raises ``NotImplementedError`` at render time rather than taking a slow path.
"""
from typing import Any, Callable, Dict, NamedTuple, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module
from torch.nn.functional import interpolate

from thre3d_atom.thre3d_reprs.constants import CONFIG_DICT, STATE_DICT, THRE3D_REPR, u_ATTN, u_DENSITIES, u_FEATURES
from voxe_b200 import _native as nat
from voxe_b200.render_function import FusedGridSpec, PackedGradAccumulator, PackedVolumeCache


class VoxelSize(NamedTuple):
    """edge lengths of one voxel (anisotropic voxels allowed)"""

    x_size: float = 1.0
    y_size: float = 1.0
    z_size: float = 1.0


class VoxelGridLocation(NamedTuple):
    """world-space position of the grid centre; the grid is always axis aligned"""

    x_coord: float = 0.0
    y_coord: float = 0.0
    z_coord: float = 0.0


class AxisAlignedBoundingBox(NamedTuple):
    x_range: Tuple[float, float]
    y_range: Tuple[float, float]
    z_range: Tuple[float, float]


def _classify_preactivation(fn) -> int:
    if isinstance(fn, torch.nn.Identity):
        return nat.PREACT_IDENTITY
    if fn is torch.abs:
        return nat.PREACT_ABS
    raise NotImplementedError(
        f"density_preactivation {fn!r} is outside the set fused into the CUDA kernels (torch.nn.Identity(), torch.abs)"
    )


def _classify_postactivation(fn) -> int:
    if isinstance(fn, torch.nn.Identity):
        return nat.POSTACT_IDENTITY
    if isinstance(fn, torch.nn.ReLU) or fn is torch.relu or fn is torch.nn.functional.relu:
        return nat.POSTACT_RELU
    if isinstance(fn, torch.nn.Softplus):
        if fn.beta != 1 or fn.threshold != 20:
            raise NotImplementedError("only torch.nn.Softplus(beta=1, threshold=20) is fused into the CUDA kernels")
        return nat.POSTACT_SOFTPLUS
    raise NotImplementedError(
        f"density_postactivation {fn!r} is outside the set fused into the CUDA kernels (Identity, ReLU, Softplus)"
    )


class VoxelGrid(Module):
    def __init__(
        self,
        # grid values:
        densities: Tensor,
        features: Tensor,
        # grid coordinate-space properties:
        voxel_size: VoxelSize,
        grid_location: Optional[VoxelGridLocation] = VoxelGridLocation(),
        # density activations:
        density_preactivation: Callable[[Tensor], Tensor] = torch.abs,
        density_postactivation: Callable[[Tensor], Tensor] = torch.nn.Identity(),
        # feature activations:
        feature_preactivation: Callable[[Tensor], Tensor] = torch.nn.Identity(),
        feature_postactivation: Callable[[Tensor], Tensor] = torch.nn.Identity(),
        # radiance function / transfer function:
        radiance_transfer_function: Callable[[Tensor, Tensor], Tensor] = None,
        expected_density_scale: float = 1.0,
        tunable: bool = False,
        attn=None,
    ):
        """densities [W, D, H, 1] and features [W, D, H, F] live on the voxel centres of an axis-aligned box of
        W x D x H voxels of ``voxel_size`` centred at ``grid_location``.  ``tunable`` wraps them (and ``attn``) as
        Parameters."""
        assert len(densities.shape) == 4 and densities.shape[-1] == 1, (
            f"densities should be of shape [W x D x H x 1] as opposed to ({densities.shape})"
        )
        assert len(features.shape) == 4, f"features should be of shape [W x D x H x F] as opposed to ({features.shape})"
        assert densities.device == features.device, "densities and features are not on the same device :("
        super().__init__()

        self._density_preactivation = density_preactivation
        self._density_postactivation = density_postactivation
        self._feature_preactivation = feature_preactivation
        self._feature_postactivation = feature_postactivation
        self._radiance_transfer_function = radiance_transfer_function
        self._grid_location = grid_location
        self._voxel_size = voxel_size
        self._expected_density_scale = expected_density_scale
        self._tunable = tunable
        self.orig_densities = densities

        wrap = torch.nn.Parameter if tunable else (lambda t: t)
        self._densities = wrap(densities)
        self._features = wrap(features)
        self.attn = wrap(attn) if attn is not None else None
        self._device = features.device

        # x: width (+ve right), y: depth (+ve inwards), z: height (+ve up)
        self.width_x, self.depth_y, self.height_z = (int(s) for s in self._features.shape[:3])
        self._aabb = self._setup_bounding_box_planes()

        # packed-volume caches for the render kernels (colour render / attention render)
        self._packed = PackedVolumeCache()
        self._packed_attn = PackedVolumeCache()
        self._grad_accumulator: Optional[PackedGradAccumulator] = None  # deferred render gradients (opt-in)
        # packed gradient volumes the backward kernels scatter into; all-zero between calls (allocated on first use)
        self._grad_scratch = PackedGradAccumulator()
        self._grad_scratch_attn = PackedGradAccumulator()

    # ------------------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------------------
    def add_attn_params(self, attn):
        self.attn = torch.nn.Parameter(attn)

    def update_orig_densities(self):
        self.orig_densities = self._densities.clone().detach()

    @property
    def densities(self) -> Tensor:
        return self._densities

    @property
    def features(self) -> Tensor:
        return self._features

    def _maybe_parameter(self, value: Tensor) -> Tensor:
        if self._tunable and not isinstance(value, torch.nn.Parameter):
            return torch.nn.Parameter(value)
        return value

    @features.setter
    def features(self, features: Tensor) -> None:
        assert features.shape == self._features.shape, "new features don't match original feature tensor's dimensions"
        self._features = self._maybe_parameter(features)

    @densities.setter
    def densities(self, densities: Tensor) -> None:
        assert densities.shape == self._densities.shape, "new densities don't match original densities tensor's dimensions"
        self._densities = self._maybe_parameter(densities)

    # ------------------------------------------------------------------------------------------------------
    # geometry
    # ------------------------------------------------------------------------------------------------------
    @property
    def aabb(self) -> AxisAlignedBoundingBox:
        return self._aabb

    @property
    def grid_dims(self) -> Tuple[int, int, int]:
        return self.width_x, self.depth_y, self.height_z

    @property
    def voxel_size(self) -> VoxelSize:
        return self._voxel_size

    @voxel_size.setter
    def voxel_size(self, voxel_size: VoxelSize) -> None:
        self._voxel_size = voxel_size

    def _setup_bounding_box_planes(self) -> AxisAlignedBoundingBox:
        ranges = []
        for count, size, centre in zip(self.grid_dims, self._voxel_size, self._grid_location):
            half_extent = (count * size) / 2
            ranges.append((centre - half_extent, centre + half_extent))
        return AxisAlignedBoundingBox(*ranges)

    def get_bounding_volume_vertices(self) -> Tensor:
        (x0, x1), (y0, y1), (z0, z1) = self._aabb
        return torch.tensor([[x, y, z] for x in (x0, x1) for y in (y0, y1) for z in (z0, z1)], dtype=torch.float32)

    def test_inside_volume(self, points: Tensor) -> Tensor:
        """[N, 3] -> [N, 1] bool: strictly inside the AABB on every axis."""
        inside = torch.ones_like(points[..., 0:1], dtype=torch.bool)
        for axis, (lo, hi) in enumerate(self._aabb):
            coord = points[..., axis : axis + 1]
            inside = inside & (coord > lo) & (coord < hi)
        return inside

    # ------------------------------------------------------------------------------------------------------
    # checkpoint dictionaries
    # ------------------------------------------------------------------------------------------------------
    def get_config_dict(self) -> Dict[str, Any]:
        return {
            "grid_location": self._grid_location,
            "density_preactivation": self._density_preactivation,
            "density_postactivation": self._density_postactivation,
            "feature_preactivation": self._feature_preactivation,
            "feature_postactivation": self._feature_postactivation,
            "radiance_transfer_function": self._radiance_transfer_function,
            "expected_density_scale": self._expected_density_scale,
            "tunable": self._tunable,
        }

    def get_save_config_dict(self) -> Dict[str, Any]:
        return {**self.get_config_dict(), "voxel_size": self._voxel_size}

    def extra_repr(self) -> str:
        return (
            f"grid_dims: {(self.width_x, self.depth_y, self.height_z)}, "
            f"feature_dims: {self._features.shape[-1]}, "
            f"voxel_size: {self._voxel_size}, "
            f"grid_location: {self._grid_location}, "
            f"tunable: {self._tunable}"
        )

    # ------------------------------------------------------------------------------------------------------
    # bridge to the fused kernels
    # ------------------------------------------------------------------------------------------------------
    # attributes a FusedGridSpec is derived from: assigning any of them drops the cached descriptions (render calls ask
    # for the description every time, and looking these up through nn.Module's attribute machinery costs more than the
    # rest of the call's Python)
    _SPEC_INPUTS = frozenset({
        "_density_preactivation", "_density_postactivation", "_feature_preactivation", "_feature_postactivation", "_aabb",
        "_expected_density_scale", "width_x", "depth_y", "height_z", "_features", "attn", "_grid_location", "_voxel_size",
    })

    def __setattr__(self, name, value):
        if name in VoxelGrid._SPEC_INPUTS:
            self.__dict__.pop("_fused_spec_cache", None)
        super().__setattr__(name, value)

    def fused_spec(self, n_features: Optional[int] = None) -> FusedGridSpec:
        """Describe this grid to the CUDA kernels; rejects activations the kernels do not fuse."""
        cache = self.__dict__.get("_fused_spec_cache")
        if cache is None:
            cache = self.__dict__["_fused_spec_cache"] = {}
        spec = cache.get(n_features)
        if spec is None:
            nf = int(self._features.shape[-1]) if n_features is None else int(n_features)
            spec = cache[n_features] = self._build_fused_spec(nf)
        return spec

    def _build_fused_spec(self, n_features: int) -> FusedGridSpec:
        for name in ("_feature_preactivation", "_feature_postactivation"):
            if not isinstance(getattr(self, name), torch.nn.Identity):
                raise NotImplementedError(f"{name[1:]} must be torch.nn.Identity() on the fused render path")
        return FusedGridSpec(
            dims=self.grid_dims,
            n_features=int(n_features),
            aabb=tuple((float(lo), float(hi)) for lo, hi in self._aabb),  # fixed at construction, as upstream
            density_scale=float(self._expected_density_scale),
            preact=_classify_preactivation(self._density_preactivation),
            postact=_classify_postactivation(self._density_postactivation),
        )

    def packed_cache(self, attn: bool = False) -> PackedVolumeCache:
        return self._packed_attn if attn else self._packed

    def invalidate_packed_cache(self) -> None:
        """Call after writing grid values through ``.data`` (which autograd's version counter does not see)."""
        self._packed.invalidate()
        self._packed_attn.invalidate()

    def accumulate_render_gradients(self, enabled: bool = True, trail: bool = False) -> None:
        """Opt into deferred gradients: render backward passes scatter into one persistent packed volume instead of
        producing dense ``.grad`` tensors per call; ``materialize_render_gradients()`` (called automatically before any
        ``torch.optim`` step once ``voxe_b200.optim`` is imported, or replaced by ``FusedVoxelAdam``) makes them visible."""
        if enabled and self._grad_accumulator is None:
            from voxe_b200 import optim  # registers the optimiser pre-step hook

            self._grad_accumulator = PackedGradAccumulator()
            optim.track_grid(self)
            if trail:  # large grids: remember which bricks a step touched; the hand-over into .grad then visits only those
                spec = self.fused_spec()
                self._grad_accumulator.enable_trail(spec, self._packed.get(spec, self._densities, self._features))
        elif enabled and trail and not self._grad_accumulator.sparse_sink:  # already deferred: start keeping the trail now
            self.materialize_render_gradients()
            spec = self.fused_spec()
            self._grad_accumulator.enable_trail(spec, self._packed.get(spec, self._densities, self._features))
        elif not enabled and self._grad_accumulator is not None:
            self.materialize_render_gradients()
            self._grad_accumulator = None

    @property
    def render_gradient_accumulator(self) -> Optional[PackedGradAccumulator]:
        return self._grad_accumulator

    def render_gradient_scratch(self, attn: bool = False) -> PackedGradAccumulator:
        return self._grad_scratch_attn if attn else self._grad_scratch

    def materialize_render_gradients(self) -> None:
        if self._grad_accumulator is not None:
            self._grad_accumulator.materialize(self.fused_spec(), self._densities, self._features)

    def forward(self, points: Tensor, viewdirs: Optional[Tensor] = None) -> Tensor:
        """Features / radiance and density at ``points`` [N, 3] -> [N, F + 1] (voxels.py:287-345 upstream): one
        ``voxe_query_points`` launch over the packed volume instead of two ``grid_sample`` calls and a full-grid
        ``densities * scale`` pass.  The renderers do not come through here -- sampling, interpolation and compositing are
        one kernel there -- this is the stand-alone query of the reference's API."""
        from voxe_b200.render_function import query_points

        out = query_points(self.fused_spec(), self._densities, self._features, points, cache=self._packed)
        return self._with_radiance(out, viewdirs)

    def forward_attn(self, points: Tensor, viewdirs: Optional[Tensor] = None, orig_densities=False) -> Tensor:
        """[N, 2] = (interpolated attention value, density) at ``points`` (voxels.py:347-406 upstream); ``orig_densities``
        reads the densities kept by ``update_orig_densities``."""
        from voxe_b200.render_function import query_points

        assert self.attn is not None, "forward_attn needs the attention grid (add_attn_params / attn=...)"
        densities = self.orig_densities if orig_densities else self._densities
        out = query_points(self.fused_spec(n_features=1), densities, self.attn, points, cache=self._packed_attn)
        return self._with_radiance(out, viewdirs)

    def _with_radiance(self, out: Tensor, viewdirs: Optional[Tensor]) -> Tensor:
        if self._radiance_transfer_function is None or viewdirs is None:
            return out
        # voxels.py:336-340: a user-supplied callable (None in every script of the reference), applied to the features
        radiance = self._radiance_transfer_function(out[:, :-1], viewdirs)
        return torch.cat([radiance, out[:, -1:]], dim=-1)


def _resample(tensor: Tensor, output_size: Tuple[int, int, int], mode: str) -> Tensor:
    """[X, Y, Z, C] -> [X2, Y2, Z2, C] as ``interpolate(mode="trilinear", align_corners=False, size=output_size)`` does.
    CUDA fp32 grids go through ``voxe_resample_grid`` (one launch per tensor, channel-last in and out: no cat / permute /
    slice copies); host-side tensors -- a grid that has not been moved to its device yet -- keep the torch call."""
    if tensor.is_cuda and tensor.dtype == torch.float32:
        if mode != "trilinear":
            raise NotImplementedError(f"mode={mode!r}: the fused rescale implements 'trilinear' (the only mode the reference's callers use)")
        import ctypes

        from voxe_b200 import _native as nat

        src = tensor.detach().contiguous()
        out = torch.empty((*output_size, src.shape[-1]), dtype=torch.float32, device=src.device)
        d_in, d_out = (ctypes.c_int32 * 3)(*src.shape[:3]), (ctypes.c_int32 * 3)(*output_size)
        with torch.cuda.device(src.device):
            nat.check(nat.load_library().voxe_resample_grid(src.data_ptr(), d_in, int(src.shape[-1]), out.data_ptr(), d_out,
                                                            torch.cuda.current_stream(src.device).cuda_stream), "voxe_resample_grid")
        return out
    return interpolate(tensor.permute(3, 0, 1, 2)[None, ...], size=output_size, mode=mode, align_corners=False,
                       recompute_scale_factor=False)[0].permute(1, 2, 3, 0)


def _rescaled(voxel_grid: VoxelGrid, unified: Tensor, output_size: Tuple[int, int, int], mode: str):
    output_size = tuple(int(v) for v in output_size)
    resized = _resample(unified, output_size, mode)
    assert resized.shape[:-1] == output_size
    old = voxel_grid.voxel_size
    new_voxel_size = VoxelSize(
        (old.x_size * voxel_grid.width_x) / output_size[0],
        (old.y_size * voxel_grid.depth_y) / output_size[1],
        (old.z_size * voxel_grid.height_z) / output_size[2],
    )
    return resized, new_voxel_size


def scale_voxel_grid_with_required_output_size(
    voxel_grid: VoxelGrid, output_size: Tuple[int, int, int], mode: str = "trilinear"
) -> VoxelGrid:
    """Resample features+densities to ``output_size`` voxels covering the same world extent (progressive training).  The
    two tensors are rescaled separately (interpolation acts per channel, so this equals the reference's concatenate ->
    interpolate -> slice) and land in the new grid as contiguous tensors of their own."""
    new_features, new_voxel_size = _rescaled(voxel_grid, voxel_grid.features, output_size, mode)
    new_densities, _ = _rescaled(voxel_grid, voxel_grid.densities, output_size, mode)
    return VoxelGrid(densities=new_densities, features=new_features, voxel_size=new_voxel_size, **voxel_grid.get_config_dict())


def scale_voxel_grid_with_required_output_size_attn(
    voxel_grid: VoxelGrid, output_size: Tuple[int, int, int], mode: str = "trilinear"
) -> VoxelGrid:
    """Variant carrying the attention channel; slices mirror the reference (voxels.py:449-488) verbatim, including its
    overlapping ``[..., -2:]`` / ``[..., :-1]`` views."""
    unified = torch.cat([voxel_grid.features, voxel_grid.densities, voxel_grid.attn], dim=-1)
    resized, new_voxel_size = _rescaled(voxel_grid, unified, output_size, mode)
    return VoxelGrid(
        densities=resized[..., -2:],
        features=resized[..., :-2],
        attn=resized[..., :-1],
        voxel_size=new_voxel_size,
        **voxel_grid.get_config_dict(),
    )


def create_voxel_grid_from_saved_info_dict(saved_info: Dict[str, Any]) -> VoxelGrid:
    state = saved_info[THRE3D_REPR][STATE_DICT]
    voxel_grid = VoxelGrid(
        densities=torch.empty_like(state[u_DENSITIES]),
        features=torch.empty_like(state[u_FEATURES]),
        **saved_info[THRE3D_REPR][CONFIG_DICT],
    )
    voxel_grid.load_state_dict(state)
    return voxel_grid


def create_voxel_grid_from_saved_info_dict_attn(saved_info: Dict[str, Any], load_attn=False) -> VoxelGrid:
    state = saved_info[THRE3D_REPR][STATE_DICT]
    densities = torch.empty_like(state[u_DENSITIES])
    features = torch.empty_like(state[u_FEATURES])
    config = saved_info[THRE3D_REPR][CONFIG_DICT]
    if load_attn:
        voxel_grid = VoxelGrid(densities=densities, features=features, attn=torch.empty_like(state[u_ATTN]), **config)
        voxel_grid.load_state_dict(state)
        return voxel_grid
    voxel_grid = VoxelGrid(densities=densities, features=features, **config)
    voxel_grid.load_state_dict(state)
    # a strongly negative logit keeps the fresh attention map near zero after the sigmoid
    voxel_grid.add_attn_params(torch.ones_like(densities) * (-20.0))
    return voxel_grid
