"""Row f2: the per-step grid passes.  ``voxe_adam_step`` against ``torch.optim.Adam`` (the optimiser the reference builds
at modules/trainers.py:252-255 and modules/sds_trainer.py:198-203), deferred gradient accumulation against the regular
autograd path, and whole training loops through both."""
import copy

import pytest
import torch

from _golden import grad_errors, load_case

pytestmark = pytest.mark.gpu


def _grid(dims=(9, 14, 11), n_feat=12, seed=0, post=None):
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    g = torch.Generator().manual_seed(seed)
    dens = (torch.rand((*dims, 1), generator=g) * 2 - 1).cuda()
    feat = (torch.rand((*dims, n_feat), generator=g) * 2 - 1).cuda()
    return VoxelGrid(dens, feat, VoxelSize(*(3.0 / d for d in dims)), density_preactivation=torch.nn.Identity(),
                     density_postactivation=post or torch.nn.Softplus(), expected_density_scale=8.0, tunable=True)


@pytest.mark.parametrize("dims,n_feat", [((9, 14, 11), 12), ((16, 16, 16), 3), ((7, 5, 3), 27), ((64, 64, 64), 3)])
def test_adam_kernel_matches_torch_adam(dims, n_feat):
    """Deterministic dense gradients, four steps, odd grid sizes (partial bricks) and every channel padding pattern."""
    from voxe_b200.optim import FusedVoxelAdam

    grid = _grid(dims, n_feat)
    ref_d = torch.nn.Parameter(grid.densities.detach().clone())
    ref_f = torch.nn.Parameter(grid.features.detach().clone())
    ref = torch.optim.Adam([{"params": [ref_d, ref_f], "lr": 0.03}], betas=(0.9, 0.999))
    ours = FusedVoxelAdam(grid, lr=0.03, betas=(0.9, 0.999))
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(4):
        gd = torch.randn(ref_d.shape, device="cuda", generator=g) * 10.0 ** (-step)
        gf = torch.randn(ref_f.shape, device="cuda", generator=g)
        gf[::2] = 0.0  # exact zeros stay exactly put
        ref_d.grad, ref_f.grad = gd.clone(), gf.clone()
        grid.densities.grad, grid.features.grad = gd.clone(), gf.clone()
        ref.step()
        ours.step()
        moments = ours.moments()
        for (a, b), (m, v) in zip(((grid.densities, ref_d), (grid.features, ref_f)), moments):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), float((a - b).abs().max())
            assert torch.allclose(m, ref.state[b]["exp_avg"], rtol=2e-6, atol=1e-9)
            assert torch.allclose(v, ref.state[b]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    # checkpoints keep torch.optim.Adam's layout and round-trip, also into a stock torch.optim.Adam
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 4.0
    assert sd["state"][1]["exp_avg"].shape == ref_f.shape
    stock = torch.optim.Adam([{"params": [ref_d, ref_f], "lr": 0.03}], betas=(0.9, 0.999))
    stock.load_state_dict(sd)
    again = FusedVoxelAdam(grid, lr=0.01)
    again.load_state_dict(sd)
    assert again.param_groups[0]["lr"] == 0.03
    gd, gf = torch.randn_like(ref_d), torch.randn_like(ref_f)
    ref_d.grad, ref_f.grad = gd.clone(), gf.clone()
    grid.densities.grad, grid.features.grad = gd.clone(), gf.clone()
    stock.step()
    again.step()
    assert torch.allclose(grid.features, ref_f, rtol=2e-6, atol=2e-7) and torch.allclose(grid.densities, ref_d, rtol=2e-6, atol=2e-7)


def _render_loss(grid, meta, a, sel=slice(None)):
    from _product import make_config
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid

    out = render_sh_voxel_grid(grid, Rays(a["rays_o"].cuda()[sel], a["rays_d"].cuda()[sel]), make_config(meta))
    return (out.colour * a["g_colour"].cuda()[sel]).sum()


def test_deferred_gradients_equal_autograd_gradients():
    from _product import make_grid

    meta, a = load_case("softplus_white")
    regular = make_grid(meta, a["densities"], a["features"], "cuda")
    deferred = make_grid(meta, a["densities"], a["features"], "cuda")
    deferred.accumulate_render_gradients()
    for sel in (slice(0, 70), slice(70, None)):  # two backward passes accumulate
        _render_loss(regular, meta, a, sel).backward()
        _render_loss(deferred, meta, a, sel).backward()
    assert deferred.densities.grad is None and deferred.features.grad is None  # nothing handed to autograd yet
    # a torch-side loss on the parameters flows through autograd as usual and is added to at materialisation
    (regular.densities ** 2).sum().backward()
    (deferred.densities ** 2).sum().backward()
    deferred.materialize_render_gradients()
    for got, want in ((deferred.densities.grad, regular.densities.grad), (deferred.features.grad, regular.features.grad)):
        l2, linf = grad_errors(got.cpu(), want.cpu())
        assert l2 <= 1e-6 and linf <= 1e-6
    # materialising twice adds nothing
    before = deferred.features.grad.clone()
    deferred.materialize_render_gradients()
    assert torch.equal(before, deferred.features.grad)


def test_deferred_gradients_with_a_brick_trail_equal_autograd_gradients():
    """`accumulate_render_gradients(trail=True)`: every backward of a step tags the bricks it scatters into and the
    hand-over visits only those -- same `.grad` as autograd, over several steps (fresh tag per step, flags never cleared),
    and the sink volume is all-zero after each hand-over."""
    from _product import make_grid

    meta, a = load_case("softplus_white")
    regular = make_grid(meta, a["densities"], a["features"], "cuda")
    deferred = make_grid(meta, a["densities"], a["features"], "cuda")
    deferred.accumulate_render_gradients(trail=True)
    acc = deferred.render_gradient_accumulator
    assert acc.sparse_sink and acc.touched is not None and int(acc.touch_tag[0]) == 1
    for step, sels in enumerate([(slice(0, 70), slice(70, None)), (slice(10, 40),), (slice(0, None),)]):
        for g in (regular, deferred):
            g.densities.grad = g.features.grad = None
        for sel in sels:
            _render_loss(regular, meta, a, sel).backward()
            _render_loss(deferred, meta, a, sel).backward()
        assert deferred.densities.grad is None and acc.dirty
        assert int((acc.touched == int(acc.touch_tag[0])).sum()) > 0  # this step's trail
        deferred.materialize_render_gradients()
        assert int(acc.touch_tag[0]) == step + 2 and not acc.dirty
        assert float(acc.buffer.abs().max()) == 0.0
        for got, want in ((deferred.densities.grad, regular.densities.grad), (deferred.features.grad, regular.features.grad)):
            l2, linf = grad_errors(got.cpu(), want.cpu())
            assert l2 <= 1e-6 and linf <= 1e-6, (step, l2, linf)


def _train(grid, optimizer, meta, a, steps, tv_weight=0.05, scheduler=None):
    losses = []
    for _ in range(steps):
        optimizer.zero_grad()
        loss = _render_loss(grid, meta, a)
        tv = grid.densities.diff(dim=0).abs().mean() * tv_weight  # torch-side regulariser (sds_trainer.py:563-567 style)
        (loss + tv).backward()
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        losses.append(float(loss))
    return losses


def test_training_loops_agree_across_the_three_step_paths():
    """(1) reference-style: autograd grads + torch Adam; (2) deferred grads + torch Adam (pre-step hook materialises);
    (3) deferred grads + FusedVoxelAdam.  Same trajectory up to float-atomic noise."""
    from _product import make_grid
    from voxe_b200.optim import FusedVoxelAdam

    meta, a = load_case("blob_softplus_all_grads_black")
    grids = [make_grid(meta, a["densities"], a["features"], "cuda") for _ in range(4)]
    grids[1].accumulate_render_gradients()
    grids[3].accumulate_render_gradients(trail=True)  # (4) as (2), the hand-over following the brick flags
    mk = lambda g: torch.optim.Adam([{"params": g.parameters(), "lr": 0.03}], betas=(0.9, 0.999))  # noqa: E731
    opts = [mk(grids[0]), mk(grids[1]), FusedVoxelAdam(grids[2], lr=0.03), mk(grids[3])]
    scheds = [torch.optim.lr_scheduler.ExponentialLR(o, gamma=0.9) for o in opts]
    runs = [_train(g, o, meta, a, 5, scheduler=s) for g, o, s in zip(grids, opts, scheds)]
    assert runs[0][-1] < runs[0][0]  # the loss <colour, G> goes down
    for other in (1, 2, 3):
        assert max(abs(x - y) for x, y in zip(runs[0], runs[other])) <= 1e-4 * max(1.0, abs(runs[0][0]))
        for p, q in ((grids[0].densities, grids[other].densities), (grids[0].features, grids[other].features)):
            # 5 steps of lr 0.03: compare against the step size, the sign-like Adam update amplifies 1e-7 gradient noise
            assert float((p - q).abs().max()) <= 5e-4, float((p - q).abs().max())
            assert float((p - q).abs().mean()) <= 1e-6
    assert abs(opts[2].param_groups[0]["lr"] - 0.03 * 0.9 ** 5) < 1e-9
    # the packed volume the fused step maintains is what a fresh pack would produce
    from voxe_b200.render_function import pack_volume

    g2 = grids[2]
    fresh = pack_volume(g2.fused_spec(), g2.densities, g2.features)
    assert torch.equal(fresh, g2.packed_cache().peek())
