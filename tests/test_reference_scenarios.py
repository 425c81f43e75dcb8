"""The scenarios of the reference's OWN tests for this path, run on the CUDA path and held to the oracle (GPU).

The reference's tests are visual smoke tests without assertions (they render and `plt.show()`):
  * thre3d_atom/thre3d_reprs/tests/test_voxels.py:88-134   `test_trilinear_interpolation_single_cube`: a 2x2x2 grid with
    voxel size 2, Identity / ReLU activations, rendered from the six axis directions (pitch +-90 included) with a 200x200
    f=240 camera, S=512, bounds (5, 18), radius 10, white background;
  * thre3d_atom/thre3d_reprs/tests/test_voxels.py:137-209  `test_render_speed`: a 128^3 grid of U(-10, 10) values with the
    constructor's DEFAULT activations (|.| pre-activation, Identity post-activation, density scale 1), 400x400 f=512, S=256,
    bounds (0.5, 8), random poses with yaw in [0, 360), pitch in [0, 180] (the camera may look from below), radius 4-5;
  * thre3d_atom/modules/tests/test_volumetric_model.py:66-103  `test_volumetric_model_render`: the same grid through
    `VolumetricModel.render(camera_pose, camera_intrinsics, verbose=True)`.
Here each scenario is rendered through the same calls and a strided subset of every frame is compared with the fp64 oracle.
The reference's configs leave `perturb_sampled_points` at its default (True), so its renders are stochastic: each scenario
is rendered twice -- as written (jittered: finite, the right shapes, close to the un-jittered picture on average) and with
`perturb_sampled_points=False`, which is held to the oracle pixel for pixel; the speed scenario also reports its time per frame."""
import time

import numpy as np
import pytest
import torch

from oracle.voxe_oracle import OracleConfig, OracleGrid, render_oracle

pytestmark = pytest.mark.gpu


def _oracle_pixels(dens, feat, ogrid, rays, sel, S, near, far):
    ocfg = OracleConfig(num_samples=S, near=near, far=far, white_bkgd=True)
    with torch.no_grad():
        return render_oracle(dens.cpu(), feat.cpu(), ogrid, rays.origins[sel].cpu(), rays.directions[sel].cpu(), ocfg, dtype=torch.float64)


def test_single_cube_from_the_six_axis_directions():
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.constants import EXTRA_ACCUMULATED_WEIGHTS
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, pose_spherical

    dev = torch.device("cuda")
    dens = torch.tensor(np.random.uniform(-10.0, 10.0, 8), dtype=torch.float32).reshape(2, 2, 2, 1)
    feat = torch.tensor([10.0, -10.0, -10.0, -10.0, 10.0, -10.0, -10.0, -10.0, 10.0, 10.0, 10.0, -10.0,
                         -10.0, 10.0, 10.0, 10.0, -10.0, 10.0, 10.0, 10.0, 10.0, -10.0, -10.0, -10.0]).reshape(2, 2, 2, 3)
    grid = VoxelGrid(densities=dens.to(dev), features=feat.to(dev), voxel_size=VoxelSize(2, 2, 2),
                     density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU())
    intr, bounds, S = CameraIntrinsics(200, 200, 240), CameraBounds(5.0, 18.0), 512
    ogrid = OracleGrid((2.0, 2.0, 2.0), density_scale=1.0, preact="identity", postact="relu")
    for yaw, pitch in ((0, 0), (90, 0), (180, 0), (270, 0), (0, -90), (0, 90)):
        rays = flatten_rays(cast_rays(intr, pose_spherical(yaw=yaw, pitch=pitch, radius=10.0), device=dev))
        with torch.no_grad():
            as_written = render_sh_voxel_grid(voxel_grid=grid, rays=rays, render_config=SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=bounds, white_bkgd=True))
            out = render_sh_voxel_grid(voxel_grid=grid, rays=rays, render_config=SHVoxGridRenderConfig(
                num_samples_per_ray=S, camera_bounds=bounds, white_bkgd=True, perturb_sampled_points=False))
        sel = torch.arange(0, 200 * 200, 3, device=dev)
        want = _oracle_pixels(dens, feat, ogrid, rays, sel, S, *bounds)
        assert out.colour.shape == (200 * 200, 3) and out.depth.shape == (200 * 200, 1)
        assert torch.isfinite(as_written.colour).all() and float((as_written.colour - out.colour).abs().mean()) < 0.02  # jitter moves edges only
        assert (out.colour[sel].cpu() - want["colour"].float()).abs().max().item() <= 1e-4, (yaw, pitch)
        assert (out.extra[EXTRA_ACCUMULATED_WEIGHTS][sel].cpu() - want["accumulated_weight"].float()).abs().max().item() <= 1e-4, (yaw, pitch)
        assert (out.depth[sel].cpu() - want["depth"].float()).abs().max().item() <= 2e-3, (yaw, pitch)  # depths up to 18
        assert float(out.extra[EXTRA_ACCUMULATED_WEIGHTS].max()) > 0.5, "the cube must be visible from every side"


def _default_grid(dev):
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    n = 128
    dens = torch.nn.init.uniform_(torch.empty((n, n, n, 1)), -10.0, 10.0)
    feat = torch.nn.init.uniform_(torch.empty((n, n, n, 3)), -10.0, 10.0)
    grid = VoxelGrid(densities=dens.to(dev), features=feat.to(dev), voxel_size=VoxelSize(2.0 / n, 2.0 / n, 2.0 / n))  # default activations
    return grid, dens, feat, OracleGrid((2.0 / n,) * 3, density_scale=1.0, preact="abs", postact="identity")


def test_render_speed_scenario_with_default_activations():
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, pose_spherical

    dev = torch.device("cuda")
    grid, dens, feat, ogrid = _default_grid(dev)
    intr, bounds, S = CameraIntrinsics(400, 400, 512.0), CameraBounds(0.5, 8.0), 256
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=bounds, white_bkgd=True)  # as written upstream: jitter on
    cfg_fixed = SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=bounds, white_bkgd=True, perturb_sampled_points=False)
    times = []
    for k in range(12):
        yaw, pitch, radius = np.random.uniform(0.0, 360.0), np.random.uniform(0.0, 180.0), np.random.uniform(4.0, 5.0)
        rays = flatten_rays(cast_rays(intr, pose_spherical(yaw=yaw, pitch=pitch, radius=radius), device=dev))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            out = render_sh_voxel_grid(voxel_grid=grid, rays=rays, render_config=cfg, parallel_points_chunk_size=None)
        torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
        assert torch.isfinite(out.colour).all() and out.colour.shape == (160000, 3)
        if k % 4 == 0:
            sel = torch.arange(k, 160000, 401, device=dev)
            want = _oracle_pixels(dens, feat, ogrid, rays, sel, S, *bounds)
            with torch.no_grad():
                fixed = render_sh_voxel_grid(voxel_grid=grid, rays=rays, render_config=cfg_fixed)
            assert (fixed.colour[sel].cpu() - want["colour"].float()).abs().max().item() <= 1e-4, (yaw, pitch, radius)
    print(f"\n[test_render_speed scenario] 400x400x256 on a 128^3 grid: {np.mean(times[2:]):.3f} ms per frame (the reference prints the same figure)")
    assert np.mean(times[2:]) < 50.0


def test_volumetric_model_render_scenario():
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.utils.constants import EXTRA_ACCUMULATED_WEIGHTS
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, pose_spherical

    dev = torch.device("cuda")
    grid, dens, feat, ogrid = _default_grid(dev)
    bounds = CameraBounds(0.5, 8.0)
    vm = VolumetricModel(thre3d_repr=grid, render_procedure=render_sh_voxel_grid,
                         render_config=SHVoxGridRenderConfig(num_samples_per_ray=256, camera_bounds=bounds, white_bkgd=True), device=dev)
    intr = CameraIntrinsics(400, 400, 512.0)
    yaw, pitch, radius = np.random.uniform(0.0, 360.0), np.random.uniform(0.0, 180.0), np.random.uniform(4.0, 5.0)
    pose = pose_spherical(yaw=yaw, pitch=pitch, radius=radius)
    out = vm.render(pose, intr, verbose=True)  # as written upstream (jittered)
    assert out.colour.shape == (400, 400, 3) and out.depth.shape == (400, 400, 1) and out.extra[EXTRA_ACCUMULATED_WEIGHTS].shape == (400, 400, 1)
    assert torch.isfinite(out.colour).all()
    fixed = vm.render(pose, intr, perturb_sampled_points=False)  # render() forwards overrides to the config (volumetric_model.py:176-183)
    rays = flatten_rays(cast_rays(intr, pose, device=dev))
    sel = torch.arange(7, 160000, 523, device=dev)
    want = _oracle_pixels(dens, feat, ogrid, rays, sel, 256, *bounds)
    assert (fixed.colour.reshape(-1, 3)[sel].cpu() - want["colour"].float()).abs().max().item() <= 1e-4
    assert float((out.colour - fixed.colour).abs().mean()) < 0.05
