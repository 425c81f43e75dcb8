"""The grid side of one SDS edit iteration (thre3d_atom/modules/sds_trainer.py:284-334) assembled from the pieces of rows
a-f2: a differentiable full-frame render whose pixel gradient is injected (the SDS loss hands dL/dpixels over through
``SpecifyGradient``, sd.py:20-34), the density-correlation and TV regularisers on the whole grid, one Adam step.

Three routes over three iterations must leave the same parameters behind:
  A  the unmodified trainer's shape: ``total_loss.backward()`` through the render node and the reference-named regulariser
     functions (autograd), ``torch.optim.Adam``;
  B  the same with the regularisers written as the reference's torch formulas (checks A's kernels in composition:
     the render node adds its gradients into ``.grad`` itself while AccumulateGrad delivers the regularisers');
  C  the fused route: deferred render gradients + ``accumulate_*`` regulariser calls + ``FusedVoxelAdam``.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0, dims=(20, 18, 22), sh_degree=1):
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds

    g = torch.Generator().manual_seed(seed)
    n_feat = 3 * (sh_degree + 1) ** 2
    dens = torch.rand((*dims, 1), generator=g) * 2 - 0.5
    feat = torch.rand((*dims, n_feat), generator=g) * 2 - 1
    grid = VoxelGrid(dens.clone().cuda(), feat.clone().cuda(), VoxelSize(*(3.0 / d for d in dims)),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=15.0, tunable=True)
    vm = VolumetricModel(grid, render_sh_voxel_grid,
                         SHVoxGridRenderConfig(num_samples_per_ray=96, camera_bounds=CameraBounds(1.5, 7.0), white_bkgd=True,
                                               perturb_sampled_points=False), device=torch.device("cuda"))
    pretrained = (dens + 0.2 * torch.randn(dens.shape, generator=g)).cuda()
    return grid, vm, pretrained


def _frames(n=3, hw=24):
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics, pose_spherical

    g = torch.Generator().manual_seed(9)
    out = []
    for k in range(n):
        rays = flatten_rays(cast_rays(CameraIntrinsics(hw, hw, 1.2 * hw), pose_spherical(40.0 * k - 30.0, -35.0, 4.0), device=torch.device("cuda")))
        out.append((rays, torch.randn(hw * hw, 3, generator=g).cuda() * 0.05))
    return out


def _tv_torch(x):
    return (x.diff(dim=0).abs().mean() + x.diff(dim=1).abs().mean() + x.diff(dim=2).abs().mean()) / 3


def _corr_torch(a, b):
    cov = (a - torch.mean(a)) * (b - torch.mean(b))
    den = torch.sqrt(torch.mean((a - torch.mean(a)) ** 2) * torch.mean((b - torch.mean(b)) ** 2))
    return 1.0 - torch.mean(cov / (den + 1e-7))


W_CORR, W_TVD, W_TVF = 200.0, 0.5, 0.05


def _run(route):
    from voxe_b200 import regularizers as reg
    from voxe_b200.optim import FusedVoxelAdam

    grid, vm, pretrained = _setup()
    if route == "C":
        opt = FusedVoxelAdam(grid, lr=0.01)
    else:
        opt = torch.optim.Adam(grid.parameters(), lr=0.01)
    losses = []
    for rays, g_pixels in _frames():
        out = vm.render_rays(rays)
        total = (out.colour * g_pixels).sum()  # SpecifyGradient: d total / d colour = the injected gradient
        d, f = grid.densities, grid.features
        if route == "A":
            corr, _ = reg.density_correlation_loss_fn(sds_density=d, regular_density=pretrained)
            total = total + corr * W_CORR + reg._tv_loss_on_grid(torch.nn.ReLU()(d)) * W_TVD + reg._tv_loss_on_grid(f) * W_TVF
        elif route == "B":
            total = total + _corr_torch(d, pretrained) * W_CORR + _tv_torch(torch.relu(d)) * W_TVD + _tv_torch(f) * W_TVF
        total.backward()
        if route == "C":
            reg.accumulate_density_loss_gradient(d, pretrained, W_CORR)
            reg.accumulate_tv_gradient(d, W_TVD, relu=True)
            reg.accumulate_tv_gradient(f, W_TVF)
        opt.step()
        opt.zero_grad()
        losses.append(float(total.detach()))
    return grid.densities.detach().clone(), grid.features.detach().clone(), losses


def test_three_routes_agree():
    ref_d, ref_f, ref_l = _run("B")
    for route in ("A", "C"):
        d, f, l = _run(route)
        # Adam normalises the step: where a gradient is a rounding error away from 0 the update direction is decided by that
        # rounding, so a handful of voxels may differ by up to 2 * lr per step; everything else agrees closely
        for got, want in ((d, ref_d), (f, ref_f)):
            diff = (got - want).abs()
            assert float(diff.median()) <= 1e-6, route
            assert float((diff > 1e-4).float().mean()) <= 2e-3, (route, float((diff > 1e-4).float().mean()))
        if route == "A":
            assert all(abs(a - b) <= 1e-4 * max(1.0, abs(b)) for a, b in zip(l, ref_l)), (l, ref_l)
