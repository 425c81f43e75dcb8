"""Helpers that drive the PRODUCT path (thre3d_atom mirror -> voxe_b200 -> libvoxe_sm100a.so) from golden / oracle inputs."""
import torch

from thre3d_atom.rendering.volumetric.render_interface import Rays
from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, _render_spec, render_sh_voxel_grid
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize
from thre3d_atom.utils.imaging_utils import CameraBounds
from voxe_b200.render_function import fused_render

ACTIVATIONS = {
    "identity": lambda: torch.nn.Identity(),
    "abs": lambda: torch.abs,
    "relu": lambda: torch.nn.ReLU(),
    "softplus": lambda: torch.nn.Softplus(),
}


def make_grid(meta, densities, features, device, tunable=True, attn=None):
    return VoxelGrid(
        densities=densities.clone().to(device),
        features=features.clone().to(device),
        voxel_size=VoxelSize(*meta["voxel_size"]),
        grid_location=VoxelGridLocation(*meta["location"]),
        density_preactivation=ACTIVATIONS[meta["preact"]](),
        density_postactivation=ACTIVATIONS[meta["postact"]](),
        expected_density_scale=meta["density_scale"],
        tunable=tunable,
        attn=None if attn is None else attn.clone().to(device),
    )


def make_config(meta):
    return SHVoxGridRenderConfig(
        num_samples_per_ray=meta["num_samples"],
        camera_bounds=CameraBounds(meta["near"], meta["far"]),
        perturb_sampled_points=meta.get("perturb", False),
        optimized_sampling=meta.get("optimized_sampling", False),
        linear_disparity_sampling=meta.get("linear_disparity_sampling", False),
        white_bkgd=meta.get("white_bkgd", False),
        render_diffuse=meta.get("render_diffuse", False),
    )


def render_case_cuda(meta, a, device="cuda"):
    """Forward + backward of one case on the GPU.  Returns a dict shaped like the oracle's."""
    grid = make_grid(meta, a["densities"], a["features"], device)
    cfg = make_config(meta)
    rays = Rays(a["rays_o"].to(device), a["rays_d"].to(device))
    if meta.get("perturb", False):  # inject the reference's jitter draws
        spec = _render_spec(cfg, grid.features.shape[-1], attn=False, per_call_sampling_flags=True)
        colour, depth, acc, disp = fused_render(
            grid.fused_spec(), spec, grid.densities, grid.features, rays.origins, rays.directions,
            cache=grid.packed_cache(), jitter=a["jitter"].to(device),
        )
    else:
        out = render_sh_voxel_grid(grid, rays, cfg)
        colour, depth, acc, disp = out.colour, out.depth, out.extra["accumulated_weight"], out.extra["disparity"]
    loss = (colour * a["g_colour"].to(device)).sum()
    if "g_depth" in a:
        loss = loss + (depth * a["g_depth"].to(device)).sum() + (acc * a["g_acc"].to(device)).sum()
        gq = a["g_disp"].to(device)
        loss = loss + (torch.where(gq != 0, disp, torch.zeros_like(disp)) * gq).sum()
    loss.backward()
    return {
        "colour": colour.detach().cpu(), "depth": depth.detach().cpu(), "accumulated_weight": acc.detach().cpu(),
        "disparity": disp.detach().cpu(), "d_densities": grid.densities.grad.cpu(), "d_features": grid.features.grad.cpu(),
    }
