"""Stratified jitter generated inside the kernels (a counter-based PCG hash keyed by the torch generator's seed / offset): the draws
are uniform, reproducible under torch.manual_seed, identical in forward and backward, and a render that uses them is
bit-identical to a render that is handed the same draws as an explicit [R,S] buffer (which is the path the reference
goldens pin)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fill(seed, offset, R, S):
    from voxe_b200 import _native as nat

    lib = nat.load_library()
    rd = nat.VoxeRenderDesc()
    rd.num_samples, rd.rng_seed, rd.rng_offset = S, seed, offset
    out = torch.empty(R, S, device="cuda")
    nat.check(lib.voxe_jitter_fill(rd, out.data_ptr(), R, torch.cuda.current_stream().cuda_stream), "voxe_jitter_fill")
    return out


def test_draws_are_uniform_and_keyed():
    u = _fill(42, 0, 4096, 256)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 2e-3 and abs(float(u.var()) - 1.0 / 12.0) < 2e-3
    hist = torch.histc(u, bins=64, min=0.0, max=1.0) / u.numel()
    assert float((hist - 1.0 / 64).abs().max()) < 1.5e-3
    # neighbouring samples / rays are uncorrelated
    assert abs(float(torch.corrcoef(torch.stack([u[:, :-1].flatten(), u[:, 1:].flatten()]))[0, 1])) < 5e-3
    assert abs(float(torch.corrcoef(torch.stack([u[:-1].flatten(), u[1:].flatten()]))[0, 1])) < 5e-3
    assert torch.equal(u, _fill(42, 0, 4096, 256))
    assert not torch.equal(u, _fill(42, 4, 4096, 256)) and not torch.equal(u, _fill(43, 0, 4096, 256))
    assert torch.equal(u[:100, :], _fill(42, 0, 100, 256))  # a draw depends on (ray, sample) only


def test_render_with_in_kernel_jitter_equals_explicit_buffer():
    from test_grad_handover import _setup
    from thre3d_atom.thre3d_reprs.renderers import _render_spec, render_sh_voxel_grid
    from voxe_b200 import render_function as rf

    grid, rays, cfg, gcol = _setup(seed=5, n_rays=1001, S=96)
    cfg.perturb_sampled_points = True
    assert not rf.STRICT_REFERENCE_RNG
    torch.manual_seed(1234)
    out = render_sh_voxel_grid(grid, rays, cfg)
    (out.colour * gcol).sum().backward()
    got = [out.colour.detach().clone(), out.depth.detach().clone(), grid.densities.grad.clone(), grid.features.grad.clone()]
    out2 = render_sh_voxel_grid(grid, rays, cfg)  # the generator advanced: another realisation
    assert not torch.equal(out2.colour, out.colour)
    torch.manual_seed(1234)
    again = render_sh_voxel_grid(grid, rays, cfg)
    assert torch.equal(again.colour, out.colour)  # reproducible under torch.manual_seed

    grid.densities.grad = grid.features.grad = None
    u = _fill(1234, 0, 1001, 96)
    spec = _render_spec(cfg, 3, attn=False, per_call_sampling_flags=True)
    colour, depth, _, _ = rf.fused_render(grid.fused_spec(), spec, grid.densities, grid.features, rays.origins, rays.directions,
                                         cache=grid.packed_cache(), jitter=u, grad_scratch=grid.render_gradient_scratch())
    (colour * gcol).sum().backward()
    assert torch.equal(colour, got[0]) and torch.equal(depth, got[1])
    for a, b in ((grid.densities.grad, got[2]), (grid.features.grad, got[3])):  # float atomics: equal up to summation order
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())
