"""How voxel gradients reach ``.grad`` (csrc/voxe_torch.cpp): the sparse direct hand-over used by plain
``loss.backward()`` must be indistinguishable from dense gradients returned to autograd, and the node must fall back to
the dense route by itself wherever the direct one would not be equivalent."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0, dims=(24, 20, 28), n_rays=777, S=64, deg=0):
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds

    g = torch.Generator().manual_seed(seed)
    dev = torch.device("cuda")
    F = 3 * (deg + 1) ** 2
    dens = (torch.rand((*dims, 1), generator=g) * 2 - 1).to(dev)
    feat = (torch.rand((*dims, F), generator=g) * 2 - 1).to(dev)
    grid = VoxelGrid(dens, feat, VoxelSize(*(3.0 / d for d in dims)), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.Softplus(), expected_density_scale=5.0, tunable=True)
    o = torch.tensor([0.0, 0.0, 4.0]) + 0.2 * torch.randn(n_rays, 3, generator=g)
    d = torch.tensor([0.0, 0.0, -1.0]) + 0.25 * torch.randn(n_rays, 3, generator=g)
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=S, camera_bounds=CameraBounds(1.5, 6.5), perturb_sampled_points=False, white_bkgd=True)
    gcol = torch.randn(n_rays, 3, generator=g).to(dev)
    return grid, Rays(o.to(dev), d.to(dev)), cfg, gcol


def _loss(grid, rays, cfg, gcol):
    from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid

    out = render_sh_voxel_grid(grid, rays, cfg)
    return (out.colour * gcol).sum() + out.depth.sum() * 0.1 + out.extra["accumulated_weight"].sum() * 0.3


def _close(a, b, tol=2e-5):
    scale = b.abs().max().item() + 1e-30
    assert (a - b).abs().max().item() <= tol * scale, ((a - b).abs().max().item(), scale)


@pytest.mark.parametrize("deg", [0, 2])
def test_direct_handover_equals_dense_gradients(deg):
    from voxe_b200 import render_function as rf

    grid, rays, cfg, gcol = _setup(deg=deg)
    assert rf.DIRECT_GRAD_ACCUMULATION
    _loss(grid, rays, cfg, gcol).backward()
    direct = [grid.densities.grad.clone(), grid.features.grad.clone()]
    assert grid.render_gradient_scratch().buffer.abs().max().item() == 0.0  # handed over and cleared
    # second call accumulates (different upstream gradient)
    _loss(grid, rays, cfg, 2.0 * gcol).backward()
    twice = [grid.densities.grad.clone(), grid.features.grad.clone()]
    grid.densities.grad = grid.features.grad = None
    rf.DIRECT_GRAD_ACCUMULATION = False
    try:
        _loss(grid, rays, cfg, gcol).backward()
        dense = [grid.densities.grad.clone(), grid.features.grad.clone()]
        _loss(grid, rays, cfg, 2.0 * gcol).backward()
        dense_twice = [grid.densities.grad.clone(), grid.features.grad.clone()]
    finally:
        rf.DIRECT_GRAD_ACCUMULATION = True
    for a, b in zip(direct + twice, dense + dense_twice):
        assert a.shape == b.shape and a.is_contiguous()
        _close(a, b)
    assert dense[1].abs().max().item() > 0


def test_autograd_grad_and_inputs_fall_back_to_dense():
    grid, rays, cfg, gcol = _setup(seed=1)
    _loss(grid, rays, cfg, gcol).backward()
    want_d, want_f = grid.densities.grad.clone(), grid.features.grad.clone()
    grid.densities.grad = grid.features.grad = None
    gd, gf = torch.autograd.grad(_loss(grid, rays, cfg, gcol), [grid.densities, grid.features])
    assert grid.densities.grad is None and grid.features.grad is None  # torch.autograd.grad must not touch .grad
    _close(gd, want_d)
    _close(gf, want_f)
    _loss(grid, rays, cfg, gcol).backward(inputs=[grid.features])
    assert grid.densities.grad is None
    _close(grid.features.grad, want_f)
    assert grid.render_gradient_scratch().buffer.abs().max().item() == 0.0


def test_hooks_force_the_dense_route_and_fire():
    grid, rays, cfg, gcol = _setup(seed=2)
    _loss(grid, rays, cfg, gcol).backward()
    want_f = grid.features.grad.clone()
    grid.densities.grad = grid.features.grad = None
    seen = []
    handle = grid.features.register_hook(lambda g: seen.append(g.shape) or g * 2.0)
    _loss(grid, rays, cfg, gcol).backward()
    handle.remove()
    assert seen == [want_f.shape]
    _close(grid.features.grad, 2.0 * want_f)


def test_other_losses_on_the_parameters_accumulate_with_the_render():
    """sds_trainer.py:305-326 adds regularisers computed on the raw grids to the render loss."""
    grid, rays, cfg, gcol = _setup(seed=3)
    _loss(grid, rays, cfg, gcol).backward()
    want_d, want_f = grid.densities.grad.clone(), grid.features.grad.clone()
    grid.densities.grad = grid.features.grad = None
    (_loss(grid, rays, cfg, gcol) + (grid.densities ** 2).sum() + grid.features.sum()).backward()
    _close(grid.densities.grad, want_d + 2.0 * grid.densities.detach())
    _close(grid.features.grad, want_f + 1.0)


def test_frozen_densities_get_no_gradient():
    grid, rays, cfg, gcol = _setup(seed=4)
    grid.densities.requires_grad_(False)
    _loss(grid, rays, cfg, gcol).backward()
    assert grid.densities.grad is None and grid.features.grad is not None
    assert grid.render_gradient_scratch().buffer.abs().max().item() == 0.0
