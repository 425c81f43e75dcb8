"""Ray casting, batching and output collation around the render call
(reference: thre3d_atom/rendering/volumetric/utils/misc.py:12-231)."""
from typing import Any, List, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from thre3d_atom.rendering.volumetric.render_interface import Rays, RenderOut, RenderOutAttn
from thre3d_atom.utils.constants import NUM_COORD_DIMENSIONS
from thre3d_atom.utils.imaging_utils import CameraIntrinsics, CameraPose


def cast_rays(camera_intrinsics: CameraIntrinsics, pose: CameraPose, device: torch.device = torch.device("cpu")) -> Rays:
    """Pixel-centre pinhole rays of one camera, shape [H, W, 3]; fp32 whatever the pose precision; directions are
    ``R @ ((x+.5-W/2)/f, -(y+.5-H/2)/f, -1)`` and are not normalised."""
    rotation, translation = pose.rotation, pose.translation
    if not (isinstance(rotation, Tensor) and isinstance(translation, Tensor)):
        rotation, translation = torch.from_numpy(rotation), torch.from_numpy(translation)
    rotation, translation = rotation.to(device), translation.to(device)

    height, width, focal = camera_intrinsics
    cols = torch.linspace(0.5, width - 0.5, width, dtype=torch.float32, device=device)
    rows = torch.linspace(0.5, height - 0.5, height, dtype=torch.float32, device=device)
    y_coords, x_coords = torch.meshgrid(rows, cols, indexing="ij")  # [H, W] each
    camera_dirs = torch.stack(
        [(x_coords - width * 0.5) / focal, -(y_coords - height * 0.5) / focal, -torch.ones_like(x_coords)], dim=-1
    )
    directions = (rotation @ camera_dirs[..., None])[..., 0]
    origins = torch.broadcast_to(translation.squeeze(), directions.shape)
    return Rays(origins, directions)


def flatten_rays(rays: Rays) -> Rays:
    return Rays(
        origins=rays.origins.reshape(-1, NUM_COORD_DIMENSIONS),
        directions=rays.directions.reshape(-1, NUM_COORD_DIMENSIONS),
    )


def collate_rays(rays_list: Sequence[Rays]) -> Rays:
    return Rays(
        origins=torch.cat([r.origins for r in rays_list], dim=0),
        directions=torch.cat([r.directions for r in rays_list], dim=0),
    )


def collate_rays_unflattened(rays_list: Sequence[Rays]) -> Rays:
    return Rays(
        origins=torch.stack([r.origins for r in rays_list], dim=0),
        directions=torch.stack([r.directions for r in rays_list], dim=0),
    )


def compute_expected_density_scale_for_relu_field_grid(grid_world_size: Tuple[float, float, float]) -> float:
    """100 * sqrt(27) / |diagonal| / 3 -- the density scale the ReLU-field scripts pass to VoxelGrid (33.33 for a 3^3 box)."""
    diagonal = float(np.sqrt(np.sum([extent**2 for extent in grid_world_size])))
    return ((float(np.sqrt(3.0**3)) * 100.0) / diagonal) / NUM_COORD_DIMENSIONS


def ndcize_rays(rays: Rays, camera_intrinsics: CameraIntrinsics) -> Rays:
    """Rays in normalised device coordinates (forward-facing scenes; the canvas becomes the cube [-1, 1]^3): origins are
    moved onto the near plane z = -1 and both origins and directions go through the pinhole projection
    (misc.py:90-123; used by visualizations/static.py:48-49)."""
    height, width, focal = camera_intrinsics
    near = 1.0
    origins, directions = rays.origins, rays.directions
    shift = -(near + origins[..., 2]) / directions[..., 2]
    origins = origins + shift[..., None] * directions
    sx, sy = -1.0 / (width / (2.0 * focal)), -1.0 / (height / (2.0 * focal))
    ox_z, oy_z = origins[..., 0] / origins[..., 2], origins[..., 1] / origins[..., 2]
    ndc_o = torch.stack([sx * ox_z, sy * oy_z, 1.0 + 2.0 * near / origins[..., 2]], -1)
    ndc_d = torch.stack(
        [sx * (directions[..., 0] / directions[..., 2] - ox_z), sy * (directions[..., 1] / directions[..., 2] - oy_z),
         -2.0 * near / origins[..., 2]], -1)
    return Rays(ndc_o, ndc_d)


def sample_random_rays_and_pixels_synchronously(rays: Rays, pixels: Tensor, sample_size: int) -> Tuple[Rays, Tensor]:
    """Random ray batch for reconstruction training: the first ``sample_size`` entries of a permutation of all pixels.
    fp32 CUDA tensors take one launch of ``voxe_sample_rays`` (distinct indices from a keyed permutation evaluated on demand,
    rays and pixels gathered in the same kernel) instead of a shuffle of all B*H*W pixels; see ``voxe_b200.sampling`` --
    its ``sample_rays_from_cameras`` also removes the per-view ``cast_rays`` this signature presupposes."""
    if pixels.is_cuda and pixels.dtype == torch.float32 and rays.origins.dtype == torch.float32 and pixels.dim() == 2:
        from voxe_b200.sampling import sample_random_rays_and_pixels

        origins, directions, picked, _ = sample_random_rays_and_pixels(rays.origins, rays.directions, pixels, sample_size)
        return Rays(origins, directions), picked
    chosen = torch.randperm(pixels.shape[0], dtype=torch.long, device=pixels.device)[:sample_size]
    return Rays(rays.origins[chosen, :], rays.directions[chosen, :]), pixels[chosen, :]


def sample_rays_and_pixels_synchronously(rays: Rays, pixels: Tensor, indices: list, sample_size: int):
    """Whole-image variant used by the attention trainer: picks ``sample_size`` images (rays [B,H,W,3], pixels [B,C,H,W])."""
    chosen = torch.randperm(pixels.shape[0], dtype=torch.long, device=pixels.device)[:sample_size]
    picked_rays = flatten_rays(Rays(rays.origins[chosen, :], rays.directions[chosen, :]))
    picked_pixels = pixels[chosen, :].permute(0, 2, 3, 1).reshape(-1, pixels.shape[1])
    picked_indices = indices[chosen.to("cpu")]
    if sample_size == 1:
        picked_indices = [picked_indices]
    return picked_rays, picked_pixels, picked_indices, chosen.tolist()


def sample_rays_directions_and_pixels_synchronously(rays: Rays, pixels: Tensor, directions: list, indices: list, sample_size: int):
    """Whole-image variant that also returns the view-direction words of the picked images (misc.py:160-182)."""
    chosen = torch.randperm(pixels.shape[0], dtype=torch.long, device=pixels.device)[:sample_size]
    picked_rays = flatten_rays(Rays(rays.origins[chosen, :], rays.directions[chosen, :]))
    picked_pixels = pixels[chosen, :].permute(0, 2, 3, 1).reshape(-1, pixels.shape[1])
    picked_directions = directions[chosen]
    picked_indices = indices[chosen.to("cpu")]
    if sample_size == 1:
        picked_directions, picked_indices = [picked_directions], [picked_indices]
    return picked_rays, picked_pixels, picked_directions, picked_indices, chosen.tolist()


def _collate(chunks: Sequence[Any], main: str, out_type):
    mains = [getattr(c, main) for c in chunks]
    depths = [c.depth for c in chunks]
    extra: dict = {}
    for c in chunks:
        for key, value in c.extra.items():
            extra.setdefault(key, []).append(value)
    return out_type(
        **{main: torch.cat(mains, dim=0)},
        depth=torch.cat(depths, dim=0),
        extra={key: torch.cat(values, dim=0) for key, values in extra.items()},
    )


def collate_rendered_output(rendered_chunks: Sequence[RenderOut]) -> RenderOut:
    return _collate(rendered_chunks, "colour", RenderOut)


def collate_rendered_output_attn(rendered_chunks: Sequence[RenderOutAttn]) -> RenderOutAttn:
    return _collate(rendered_chunks, "attn", RenderOutAttn)


def _as_image(out, main: str, camera_intrinsics: CameraIntrinsics):
    shape = (camera_intrinsics.height, camera_intrinsics.width, -1)
    return type(out)(
        **{main: getattr(out, main).reshape(*shape)},
        depth=out.depth.reshape(*shape),
        extra={key: value.reshape(*shape) for key, value in out.extra.items()},
    )


def reshape_rendered_output(rendered_output: RenderOut, camera_intrinsics: CameraIntrinsics) -> RenderOut:
    return _as_image(rendered_output, "colour", camera_intrinsics)


def reshape_rendered_output_attn(rendered_output: RenderOutAttn, camera_intrinsics: CameraIntrinsics) -> RenderOutAttn:
    return _as_image(rendered_output, "attn", camera_intrinsics)
