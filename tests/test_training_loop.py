"""The reference's reconstruction loop, end to end on the GPU path (thre3d_atom/modules/trainers.py:281-351): a target scene is
rendered to posed images, a fresh grid is trained on random ray batches of those images -- `sample_random_rays_and_pixels_
synchronously` -> `vol_mod.render_rays` -> `l1_loss` -> `zero_grad / backward / step`, stratified jitter on -- and must fit
them.  Run once with `torch.optim.Adam` exactly as the reference builds it (trainers.py:247-255) and once with
`FusedVoxelAdam`; then the coarse grid is rescaled with `scale_voxel_grid_with_required_output_size` (the stage change of
trainers.py:481) and must keep rendering the same images."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene():
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, collate_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, pose_spherical

    dev = torch.device("cuda")
    dims = (24, 24, 24)
    ax = torch.linspace(-1, 1, dims[0])
    gx, gy, gz = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = torch.sqrt(gx**2 + gy**2 + gz**2)
    dens = (0.6 - r)[..., None].clone()                      # a ball of radius 0.6 (ReLU field: negative outside)
    feat = torch.stack([gx, gy * 0.5 + 0.3, -gz], dim=-1) * 3.0  # colour varies over the ball
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=96, camera_bounds=CameraBounds(2.0, 6.0), white_bkgd=True, perturb_sampled_points=True)

    def model(d, f):
        grid = VoxelGrid(d.to(dev), f.to(dev), VoxelSize(*(3.0 / n for n in dims)), density_preactivation=torch.nn.Identity(),
                         density_postactivation=torch.nn.ReLU(), expected_density_scale=33.333, tunable=True)
        return VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)

    target = model(dens, feat)
    intr = CameraIntrinsics(40, 40, 55.0)
    poses = [pose_spherical(yaw, -30.0, 4.0) for yaw in (0.0, 72.0, 144.0, 216.0, 288.0)]
    with torch.no_grad():
        images = [target.render(p, intr, perturb_sampled_points=False).colour for p in poses]          # [H, W, 3] each
    rays = collate_rays([flatten_rays(cast_rays(intr, p, device=dev)) for p in poses])
    pixels = torch.cat([im.reshape(-1, 3) for im in images]).to(dev)
    return model, dims, intr, poses, images, rays, pixels


def _train(vol_mod, optimizer, rays, pixels, iterations, batch=1024):
    from thre3d_atom.rendering.volumetric.utils.misc import sample_random_rays_and_pixels_synchronously

    losses = []
    for _ in range(iterations):
        rays_batch, pixels_batch = sample_random_rays_and_pixels_synchronously(rays, pixels, batch)
        out = vol_mod.render_rays(rays_batch)
        loss = torch.nn.functional.l1_loss(out.colour, pixels_batch)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        losses.append(float(loss.detach()))
    return losses


@pytest.mark.parametrize("which", ["torch_adam", "fused_adam"])
def test_reconstruction_loop_fits_the_posed_images(which):
    from thre3d_atom.thre3d_reprs.voxels import scale_voxel_grid_with_required_output_size
    from voxe_b200.optim import FusedVoxelAdam

    torch.manual_seed(0)
    model, dims, intr, poses, images, rays, pixels = _scene()
    g = torch.Generator().manual_seed(1)
    vol_mod = model(torch.rand((*dims, 1), generator=g) * 0.02 - 0.01, torch.rand((*dims, 3), generator=g) * 0.2 - 0.1)  # an (almost) empty grid
    grid = vol_mod.thre3d_repr
    if which == "torch_adam":
        optimizer = torch.optim.Adam(params=[{"params": grid.parameters(), "lr": 0.03}], betas=(0.9, 0.999))
    else:
        optimizer = FusedVoxelAdam(grid, lr=0.03)
    losses = _train(vol_mod, optimizer, rays, pixels, iterations=300)
    first, last = sum(losses[:10]) / 10, sum(losses[-10:]) / 10
    assert last < 0.25 * first, f"{which}: L1 {first:.4f} -> {last:.4f}"
    with torch.no_grad():
        psnr = []
        for pose, want in zip(poses, images):
            got = vol_mod.render(pose, intr, perturb_sampled_points=False).colour
            psnr.append(float(-10.0 * torch.log10(torch.mean((got - want) ** 2))))
    assert min(psnr) > 20.0, f"{which}: PSNR per view {psnr}"

    # stage change: the trained coarse grid resampled to 1.5x the resolution renders the same pictures
    vol_mod.thre3d_repr = scale_voxel_grid_with_required_output_size(vol_mod.thre3d_repr, tuple(int(n * 1.5) for n in dims))
    with torch.no_grad():
        got = vol_mod.render(poses[0], intr, perturb_sampled_points=False).colour
    assert float(-10.0 * torch.log10(torch.mean((got - images[0]) ** 2))) > 18.0
