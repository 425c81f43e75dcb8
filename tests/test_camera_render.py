"""Whole-camera inference kernel (voxe_render_camera: rays generated in-kernel, one thread per pixel, early termination)
against the training kernels driven with cast_rays() tensors, which the reference goldens pin."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(deg=0, S=128, perturb=False, white=True, optimized=False, postact="softplus"):
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds

    g = torch.Generator().manual_seed(11)
    dims = (40, 36, 44)
    xs = [torch.linspace(-1, 1, d) for d in dims]
    r = torch.sqrt(xs[0][:, None, None] ** 2 + xs[1][None, :, None] ** 2 + xs[2][None, None, :] ** 2)
    dens = (0.35 * (1.0 - r))[..., None] + 0.02 * torch.randn((*dims, 1), generator=g)   # semi-transparent blob
    feat = torch.rand((*dims, 3 * (deg + 1) ** 2), generator=g) * 2 - 1
    act = torch.nn.Softplus() if postact == "softplus" else torch.nn.ReLU()
    grid = VoxelGrid(dens.cuda(), feat.cuda(), VoxelSize(*(3.0 / d for d in dims)), density_preactivation=torch.nn.Identity(),
                     density_postactivation=act, expected_density_scale=33.333, tunable=True)
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=CameraBounds(1.8, 6.6), perturb_sampled_points=perturb,
                                white_bkgd=white, optimized_sampling=optimized)
    return VolumetricModel(grid, render_sh_voxel_grid, cfg, device=torch.device("cuda"))


def _rays_route(vm, pose, cam, **kw):
    """The chunked route of VolumetricModel.render: cast_rays tensors through the training kernels."""
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays

    rays = flatten_rays(cast_rays(cam, pose, device=torch.device("cuda")))
    with torch.no_grad():
        out = vm.render_rays(rays, **kw)
    return out


@pytest.mark.parametrize("deg,optimized,postact", [(0, False, "softplus"), (0, True, "relu"), (2, False, "softplus"), (1, True, "softplus")])
def test_camera_kernel_matches_the_ray_tensor_route(deg, optimized, postact):
    import voxe_b200.render_function as rf
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics, pose_spherical

    vm = _model(deg=deg, optimized=optimized, postact=postact)
    cam, pose = CameraIntrinsics(75, 93, 110.0), pose_spherical(33.0, -50.0, 4.0311)
    want = _rays_route(vm, pose, cam)
    saved = rf.INFERENCE_MIN_TRANSMITTANCE
    try:
        rf.INFERENCE_MIN_TRANSMITTANCE = 0.0   # every sample: only the in-kernel ray arithmetic differs (rounding)
        exact = vm.render(pose, cam)
        rf.INFERENCE_MIN_TRANSMITTANCE = 1e-5
        fast = vm.render(pose, cam, gpu_render=False)
    finally:
        rf.INFERENCE_MIN_TRANSMITTANCE = saved
    assert exact.colour.shape == (75, 93, 3) and fast.colour.device.type == "cpu"
    for got, tol_c in ((exact, 1e-5), (fast, 3e-5)):
        assert (got.colour.reshape(-1, 3).cuda() - want.colour).abs().max().item() <= tol_c
        assert (got.depth.reshape(-1, 1).cuda() - want.depth).abs().max().item() <= 2e-4
        assert (got.extra["accumulated_weight"].reshape(-1, 1).cuda() - want.extra["accumulated_weight"]).abs().max().item() <= 3e-5
    hit = want.extra["accumulated_weight"][:, 0] > 1e-3  # disparity is NaN where a ray saw nothing, in both routes
    d_got, d_want = exact.extra["disparity"].reshape(-1).cuda(), want.extra["disparity"][:, 0]
    assert torch.equal(torch.isnan(d_got), torch.isnan(d_want))
    assert ((d_got[hit] - d_want[hit]).abs() / d_want[hit].abs()).max().item() <= 1e-3


def test_pixel_ranges_tile_the_image_and_jitter_is_reproducible():
    import voxe_b200.render_function as rf
    from thre3d_atom.thre3d_reprs.renderers import _render_spec
    from thre3d_atom.utils.imaging_utils import pose_spherical

    vm = _model(perturb=True, S=64)
    grid, pose = vm.thre3d_repr, pose_spherical(10.0, -60.0, 4.0311)
    spec = _render_spec(vm.render_config, 3, attn=False, per_call_sampling_flags=True)
    args = (grid.fused_spec(), spec, grid.densities, grid.features, 48, 52, 70.0, pose.rotation, pose.translation)
    torch.manual_seed(5)
    whole = rf.fused_render_camera(*args, cache=grid.packed_cache())
    torch.manual_seed(5)
    again = rf.fused_render_camera(*args, cache=grid.packed_cache())
    other = rf.fused_render_camera(*args, cache=grid.packed_cache())   # generator advanced: another jitter realisation
    assert torch.equal(whole[0], again[0]) and not torch.equal(whole[0], other[0])
    # without jitter a sub-range of pixels is exactly the corresponding slice of the whole image
    vm2 = _model(perturb=False, S=64)
    g2 = vm2.thre3d_repr
    spec2 = _render_spec(vm2.render_config, 3, attn=False, per_call_sampling_flags=True)
    args2 = (g2.fused_spec(), spec2, g2.densities, g2.features, 48, 52, 70.0, pose.rotation, pose.translation)
    full = rf.fused_render_camera(*args2, cache=g2.packed_cache())
    part = rf.fused_render_camera(*args2, cache=g2.packed_cache(), first_pixel=1000, num_pixels=777)
    for f, q in zip(full, part):
        assert torch.equal(f[1000:1777], q)
    with pytest.raises(Exception):
        rf.fused_render_camera(*args2, cache=g2.packed_cache(), first_pixel=48 * 52 - 10, num_pixels=11)


@pytest.mark.parametrize("overrides,hw", [
    (dict(num_samples_per_ray=2), (9, 7)),
    (dict(linear_disparity_sampling=True), (31, 29)),
    (dict(render_diffuse=True, white_bkgd=False), (31, 29)),
    (dict(num_samples_per_ray=513), (1, 1)),
    (dict(num_samples_per_ray=1024, white_bkgd=False), (17, 40)),
])
def test_camera_kernel_config_corners(overrides, hw):
    """Every sample evaluated (no early termination): sampling modes, tiny / odd sample counts, 1x1 images."""
    import voxe_b200.render_function as rf
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics, pose_spherical

    vm = _model(deg=1)
    cam, pose = CameraIntrinsics(hw[0], hw[1], 40.0), pose_spherical(-120.0, -35.0, 4.0311)
    want = _rays_route(vm, pose, cam, **overrides)
    saved = rf.INFERENCE_MIN_TRANSMITTANCE
    try:
        rf.INFERENCE_MIN_TRANSMITTANCE = 0.0
        got = vm.render(pose, cam, **overrides)
    finally:
        rf.INFERENCE_MIN_TRANSMITTANCE = saved
    assert got.colour.shape == (hw[0], hw[1], 3)
    assert (got.colour.reshape(-1, 3) - want.colour).abs().max().item() <= 2e-5
    assert (got.extra["accumulated_weight"].reshape(-1, 1) - want.extra["accumulated_weight"]).abs().max().item() <= 2e-5
    assert (got.depth.reshape(-1, 1) - want.depth).abs().max().item() <= 3e-4
