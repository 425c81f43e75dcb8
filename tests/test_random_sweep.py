"""Seeded random sweep of the render path against the fp64 oracle (GPU): configurations the hand-picked matrix of
tests/test_cuda_parity.py does not enumerate -- odd, tiny and anisotropic grids (down to a single voxel on an axis), S from
2 up, 1..700 rays, every SH degree, every activation pair, every sampling mode and their combinations, off-centre grids,
rays that start inside, graze or miss the box, all four upstream gradients.  Each case goes through the reference-facing
API (`render_sh_voxel_grid` + autograd); launch shapes are the library's own choice, plus a few forced odd ones
(rays per CTA not a power of two).  Tolerances as in tests/test_cuda_parity.py."""
import numpy as np
import pytest
import torch

from _golden import grad_errors
from oracle.voxe_oracle import OracleConfig, OracleGrid, relu_kink_voxels, render_oracle, render_oracle_with_grads

pytestmark = pytest.mark.gpu

ACT = {"identity": lambda: torch.nn.Identity(), "abs": lambda: torch.abs, "relu": lambda: torch.nn.ReLU(), "softplus": lambda: torch.nn.Softplus()}


def _case(seed):
    rng = np.random.default_rng(seed)
    deg = int(rng.integers(0, 4))
    dims = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 8, 13, 21, 32], size=3))
    S = int(rng.choice([2, 3, 7, 16, 33, 64, 130, 257]))
    R = int(rng.choice([1, 2, 31, 33, 100, 257, 700]))
    voxel = tuple(float(v) for v in rng.uniform(0.05, 0.4, size=3))
    location = tuple(float(v) for v in rng.uniform(-0.5, 0.5, size=3))
    pre = str(rng.choice(["identity", "abs"]))
    post = str(rng.choice(["identity", "relu", "softplus"]))
    meta = dict(seed=seed, deg=deg, dims=dims, S=S, R=R, voxel=voxel, location=location, pre=pre, post=post,
                perturb=bool(rng.integers(0, 2)), optimized=bool(rng.integers(0, 2)), disparity=bool(rng.integers(0, 2)),
                white=bool(rng.integers(0, 2)), diffuse=bool(rng.integers(0, 2)), scale=float(rng.choice([1.0, 5.0, 33.333])),
                all_grads=bool(rng.integers(0, 2)), tuning=[None, None, None, (5, 3, 128), (16, 7, 96), (8, 5, 64)][int(rng.integers(0, 6))])
    extent = np.asarray(dims) * np.asarray(voxel)
    g = torch.Generator().manual_seed(seed)
    dens = torch.rand((*dims, 1), generator=g) * 2 - (0.6 if post == "identity" else 1.0)
    if post == "identity" and pre == "identity":
        dens = dens.abs() * 0.5  # a negative density through Identity/Identity blows exp() up; upstream never trains there
    feat = torch.rand((*dims, 3 * (deg + 1) ** 2), generator=g) * 2 - 1
    # rays: origins on a sphere around the grid (some inside it), aimed at random points of a box twice the grid's size
    centre = torch.tensor(location, dtype=torch.float32)
    radius = float(np.linalg.norm(extent)) * float(rng.uniform(0.2, 1.6)) + 0.3
    dirs = torch.randn(R, 3, generator=g)
    o = centre + radius * dirs / dirs.norm(dim=-1, keepdim=True)
    spread = torch.full((R, 1), 2.0)
    spread[: max(1, R // 2)] = 0.9  # half of the rays (at least one) are aimed into the grid itself
    target = centre + (torch.rand(R, 3, generator=g) - 0.5) * spread * torch.tensor(extent, dtype=torch.float32)
    k = float(rng.uniform(0.5, 2.0))
    d = (target - o) * k  # directions are not normalised upstream either: depth 1/k reaches the target point
    if R > 4:
        d[1, 0] = 0.0  # axis-parallel component (d_a == 0)
        d[2] = torch.tensor([0.0, 0.0, -1.0]) * float(d[2].norm())  # axis-parallel ray, most likely a miss
    near, far = 0.05 / k, float(rng.uniform(1.3, 2.2)) / k  # [near, far] spans the grid for most rays, ends inside it for some
    meta.update(near=near, far=far)
    gcol = torch.randn(R, 3, generator=g)
    extra = (torch.randn(R, 1, generator=g) * 0.1, torch.randn(R, 1, generator=g) * 0.1, torch.randn(R, 1, generator=g) * 1e-3) if meta["all_grads"] else None
    jitter = torch.rand(R, S, generator=g) if meta["perturb"] else None
    return meta, dens, feat, o.contiguous(), d.contiguous(), gcol, extra, jitter


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_matches_the_oracle(seed):
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, _render_spec
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds
    from voxe_b200 import _native as nat
    from voxe_b200.render_function import fused_render

    m, dens, feat, o, d, gcol, extra, jitter = _case(seed)
    ogrid = OracleGrid(m["voxel"], location=m["location"], density_scale=m["scale"], preact=m["pre"], postact=m["post"])
    ocfg = OracleConfig(num_samples=m["S"], near=m["near"], far=m["far"], perturb=m["perturb"], optimized_sampling=m["optimized"],
                        linear_disparity_sampling=m["disparity"], white_bkgd=m["white"], render_diffuse=m["diffuse"])
    if extra is not None:
        # a ray that sees nothing has NaN disparity (accumulate.py:85-88 upstream) and would poison the scalar loss: the
        # disparity gradient is only exercised on draws where every ray hits something
        with torch.no_grad():
            probe = render_oracle(dens, feat, ogrid, o, d, ocfg, jitter=jitter, dtype=torch.float64)
        if torch.isnan(probe["disparity"]).any():
            extra = (extra[0], extra[1], None)
    want = render_oracle_with_grads(dens, feat, ogrid, o, d, ocfg, gcol, *(extra or (None, None, None)), jitter=jitter, dtype=torch.float64)
    if not all(torch.isfinite(want[k]).all() for k in ("colour", "depth", "d_densities", "d_features")):
        pytest.skip("the reference arithmetic itself overflows on this draw")

    dev = torch.device("cuda")
    grid = VoxelGrid(dens.clone().to(dev), feat.clone().to(dev), VoxelSize(*m["voxel"]), VoxelGridLocation(*m["location"]),
                     density_preactivation=ACT[m["pre"]](), density_postactivation=ACT[m["post"]](), expected_density_scale=m["scale"], tunable=True)
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=m["S"], camera_bounds=CameraBounds(m["near"], m["far"]), perturb_sampled_points=m["perturb"],
                                optimized_sampling=m["optimized"], linear_disparity_sampling=m["disparity"], white_bkgd=m["white"],
                                render_diffuse=m["diffuse"])
    spec = _render_spec(cfg, grid.features.shape[-1], attn=False, per_call_sampling_flags=True)
    try:
        if m["tuning"]:
            nat.set_tuning(*m["tuning"])
        colour, depth, acc, disp = fused_render(grid.fused_spec(), spec, grid.densities, grid.features, o.to(dev), d.to(dev),
                                                cache=grid.packed_cache(), jitter=None if jitter is None else jitter.to(dev),
                                                grad_scratch=grid.render_gradient_scratch())
        loss = (colour * gcol.to(dev)).sum()
        if extra is not None:
            loss = loss + (depth * extra[0].to(dev)).sum() + (acc * extra[1].to(dev)).sum()
            if extra[2] is not None:
                loss = loss + (disp * extra[2].to(dev)).sum()
        loss.backward()
    finally:
        nat.set_tuning(0, 0, 0)
    what = f"seed {seed}: {m}"
    assert (colour.detach().cpu() - want["colour"].float()).abs().max().item() <= 1e-4, what
    scale_z = max(1.0, m["far"])
    assert (depth.detach().cpu() - want["depth"].float()).abs().max().item() <= 2e-4 * scale_z, what
    assert (acc.detach().cpu() - want["accumulated_weight"].float()).abs().max().item() <= 1e-4, what
    assert torch.equal(torch.isnan(disp.detach().cpu()), torch.isnan(want["disparity"])), what
    gd, gf = grid.densities.grad, grid.features.grad
    gd = torch.zeros_like(grid.densities) if gd is None else gd
    gf = torch.zeros_like(grid.features) if gf is None else gf
    keep = torch.ones(m["dims"], dtype=torch.bool)
    if m["post"] == "relu":  # voxels fed by a sample sitting on the kink (derivative decided by rounding)
        keep = ~relu_kink_voxels(dens, ogrid, o, d, ocfg, jitter=jitter)
    # dL/d(density) is density_scale x delta x (T q - sum w q).  With coarse sampling (S = 16) and the ReLU-field density
    # scale, sigma x delta reaches 30 per sample: rays turn opaque within a sample or two, the two terms cancel to a fraction
    # of a percent of their size, and what is left carries the fp32 rounding of the terms -- here mostly the input rounding of
    # ex2.approx (|x| x 6e-8 relative; invisible at the benchmark's sigma x delta < 1).  The bar for d_densities is therefore
    # 2e-4 of max(its own largest entry, 3 % of what the un-cancelled terms could produce ~ density_scale x ||dL/d(features)||inf).
    term_floor = 3e-2 * m["scale"] * float(want["d_features"].abs().max())
    for name, got, ref in (("d_densities", gd.cpu() * keep[..., None], want["d_densities"] * keep[..., None]), ("d_features", gf.cpu(), want["d_features"])):
        if float(ref.abs().max()) == 0.0:
            assert float(got.abs().max()) <= 1e-12, f"{what} {name}: expected no gradient"
            continue
        l2, linf = grad_errors(got, ref)
        if name == "d_densities":
            shrink = min(1.0, float(ref.abs().max()) / max(term_floor, 1e-30))  # < 1 only when the gradient is cancellation residue
            l2, linf = l2 * shrink, linf * shrink
        assert l2 <= 2e-4 and linf <= 2e-4, f"{what} {name}: relL2 {l2:.2e} maxabs/inf {linf:.2e} (peak {float(ref.abs().max()):.3e})"
