"""Stand-alone point queries ``VoxelGrid.forward`` / ``forward_attn`` (thre3d_reprs/voxels.py:287-345, 347-406 upstream):
the oracle against goldens from the executed reference (CPU); the CUDA kernels behind ``voxe_query_points`` /
``voxe_query_points_bwd`` against both, through the reference's own call ``grid(points)`` and autograd (GPU).
Tolerances: rows 1e-5 x max(1, |row|inf) (fp32 interpolation on both sides); gradients rel-L2 / max-abs 2e-5, with ReLU
points that sit on the kink left out (their derivative is decided by rounding)."""
import numpy as np
import pytest
import torch

from _golden import grad_errors, load_npz
from oracle.voxe_oracle import OracleGrid, query_points_oracle

ACT = {"identity": lambda: torch.nn.Identity(), "abs": lambda: torch.abs, "relu": lambda: torch.nn.ReLU(), "softplus": lambda: torch.nn.Softplus()}


def _oracle_inputs(name, m, a):
    grid = OracleGrid(tuple(m["voxel"]), location=tuple(m["loc"]), density_scale=m["scale"], preact=m["pre"], postact=m["post"])
    dens = a[f"{name}/orig_densities"] if m["orig"] else a[f"{name}/densities"]
    feat = a[f"{name}/attn"] if m["attn"] else a[f"{name}/features"]
    return grid, dens, feat


def _off_kink(name, m, a, margin=1e-5):
    """Points whose interpolated pre-activated density is not within rounding of the ReLU kink."""
    if m["post"] != "relu":
        return torch.ones(m["n_points"], dtype=torch.bool)
    grid, dens, feat = _oracle_inputs(name, m, a)
    raw = query_points_oracle(dens, feat, OracleGrid(grid.voxel_size, grid.location, grid.density_scale, grid.preact, "identity"), a[f"{name}/points"])
    return raw[:, -1].abs() > margin


def test_oracle_matches_the_executed_reference():
    meta, a = load_npz("points")
    assert len(meta) >= 8
    for name, m in meta.items():
        grid, dens, feat = _oracle_inputs(name, m, a)
        keep = _off_kink(name, m, a)
        dens, feat = dens.clone().double().requires_grad_(True), feat.clone().double().requires_grad_(True)
        out = query_points_oracle(dens, feat, grid, a[f"{name}/points"])
        want = a[f"{name}/out"]
        assert out.shape == want.shape == (m["n_points"], (1 if m["attn"] else m["n_feat"]) + 1)
        assert float((out.detach() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max())), name
        # points far outside the box: interpolated values are exactly zero, the density is post(0)
        assert float(out.detach()[-3:, :-1].abs().max()) == 0.0 and float((want[-3:] - out.detach()[-3:]).abs().max()) <= 1e-7
        g = a[f"{name}/g_out"].double() * keep[:, None]
        (out * g).sum().backward()
        if bool(keep.all()):  # the stored gradients include every point
            for got, key in ((dens.grad, "d_densities"), (feat.grad, "d_attn" if m["attn"] else "d_features")):
                l2, linf = grad_errors(got, a[f"{name}/{key}"])
                assert l2 <= 2e-5 and linf <= 2e-5, (name, key, l2, linf)


def test_host_side_tensors_are_refused():
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    grid = VoxelGrid(torch.zeros(2, 2, 2, 1), torch.zeros(2, 2, 2, 3), VoxelSize(1, 1, 1))
    with pytest.raises(RuntimeError, match="CUDA only"):
        grid(torch.zeros(4, 3))


def _cuda_grid(name, m, a, requires_grad=True):
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize

    dev = torch.device("cuda")
    kw = {}
    if m["attn"]:
        kw["attn"] = a[f"{name}/attn"].to(dev)
    grid = VoxelGrid(a[f"{name}/densities"].to(dev), a[f"{name}/features"].to(dev), VoxelSize(*m["voxel"]), VoxelGridLocation(*m["loc"]),
                     density_preactivation=ACT[m["pre"]](), density_postactivation=ACT[m["post"]](), expected_density_scale=m["scale"],
                     tunable=requires_grad, **kw)
    if m["orig"]:
        grid.orig_densities = a[f"{name}/orig_densities"].to(dev).requires_grad_(requires_grad)
    return grid


@pytest.mark.gpu
def test_kernels_match_golden_and_oracle():
    from voxe_b200 import _native as nat

    meta, a = load_npz("points")
    for name, m in meta.items():
        grid = _cuda_grid(name, m, a)
        pts = a[f"{name}/points"].cuda()
        before = nat.launch_count()
        out = grid.forward_attn(pts, orig_densities=m["orig"]) if m["attn"] else grid(pts)
        assert nat.launch_count() - before == 2, "one pack + one query launch, no grid_sample"
        want = a[f"{name}/out"]
        assert out.shape == want.shape
        assert float((out.detach().cpu() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max())), name
        ogrid, odens, ofeat = _oracle_inputs(name, m, a)
        truth = query_points_oracle(odens, ofeat, ogrid, a[f"{name}/points"])
        assert float((out.detach().cpu().double() - truth).abs().max()) <= 1e-5 * max(1.0, float(truth.abs().max())), name
        assert float(out.detach()[-3:, :-1].abs().max()) == 0.0  # far outside: exact zeros

        keep = _off_kink(name, m, a)
        (out * (a[f"{name}/g_out"] * keep[:, None]).cuda()).sum().backward()
        dsrc = grid.orig_densities if m["orig"] else grid.densities
        fsrc = grid.attn if m["attn"] else grid.features
        if bool(keep.all()):
            wants = (a[f"{name}/d_densities"], a[f"{name}/d_attn" if m["attn"] else f"{name}/d_features"])
        else:  # kink points dropped on both sides: the oracle's gradients of the same masked loss
            od, of = odens.clone().double().requires_grad_(True), ofeat.clone().double().requires_grad_(True)
            (query_points_oracle(od, of, ogrid, a[f"{name}/points"]) * (a[f"{name}/g_out"].double() * keep[:, None])).sum().backward()
            wants = (od.grad, of.grad)
        for got, ref, key in ((dsrc.grad, wants[0], "d_densities"), (fsrc.grad, wants[1], "d_features")):
            l2, linf = grad_errors(got.cpu(), ref)
            assert l2 <= 2e-5 and linf <= 2e-5, (name, key, l2, linf)
        if m["attn"]:
            assert grid.features.grad is None  # the colour features take no part in forward_attn


@pytest.mark.gpu
def test_query_agrees_with_grid_sample_on_a_training_sized_grid():
    """160^3 SH-0 grid, 2^20 points (a 4096-ray x 256-sample batch's worth): against torch's own grid_sample on the GPU,
    written as the reference writes it."""
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(3)
    dens = (torch.randn((160, 160, 160, 1), generator=g) * 0.01).to(dev)
    feat = torch.randn((160, 160, 160, 3), generator=g).to(dev)
    grid = VoxelGrid(dens, feat, VoxelSize(*(3.0 / 160,) * 3), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=33.333, tunable=True)
    pts = ((torch.rand((1 << 20, 3), generator=g) - 0.5) * 3.3).to(dev)
    out = grid(pts)
    n = pts / 1.5  # AABB = [-1.5, 1.5]^3
    def sample(vol):
        return torch.nn.functional.grid_sample(vol[None].permute(0, 4, 3, 2, 1), n[None, None, None], align_corners=False).permute(0, 2, 3, 4, 1).squeeze()
    want = torch.cat([sample(feat), torch.relu(sample(dens * 33.333))[..., None]], dim=-1)
    # Both sides interpolate in fp32 from fp32 voxel coordinates u ~ 160 (ulp 1.5e-5) of a white-noise field whose neighbouring
    # voxels differ by up to ~8: the two roundings of u alone move a value by up to ~1e-4.  The fp64 oracle on a subsample
    # arbitrates: this kernel must be no further from the truth than the stock kernel is.
    assert float((out.detach() - want).abs().max()) <= 3e-4
    sub = torch.arange(0, 1 << 20, 53, device=dev)
    truth = query_points_oracle(dens.cpu(), feat.cpu(), OracleGrid((3.0 / 160,) * 3, density_scale=33.333, preact="identity", postact="relu"), pts[sub].cpu())
    err_ours, err_stock = float((out.detach()[sub].cpu().double() - truth).abs().max()), float((want[sub].cpu().double() - truth).abs().max())
    assert err_ours <= 1e-4 and err_ours <= 2.0 * err_stock + 1e-6, (err_ours, err_stock)

    def timed(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / reps

    with torch.no_grad():
        ours, stock = timed(lambda: grid(pts)), timed(lambda: torch.cat([sample(feat), torch.relu(sample(dens * 33.333))[..., None]], dim=-1))
    print(f"\n[point query] 2^20 points on 160^3 SH-0: voxe_query_points {ours * 1e3:.0f} us, grid_sample x2 as upstream {stock * 1e3:.0f} us; error against fp64: {err_ours:.1e} (stock {err_stock:.1e})")
    gout = torch.randn(out.shape, generator=g).to(dev)
    (out * gout).sum().backward()
    f2 = feat.clone().requires_grad_(True)
    (sample(f2) * gout[:, :3]).sum().backward()
    l2, linf = grad_errors(grid.features.grad.cpu(), f2.grad.cpu())
    assert l2 <= 2e-5 and linf <= 1e-4, (l2, linf)  # same coordinate rounding, seen through the trilinear weights


@pytest.mark.gpu
def test_edge_cases_empty_single_and_frozen():
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize

    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(9)
    dens, feat = torch.randn((3, 4, 5, 1), generator=g).to(dev), torch.randn((3, 4, 5, 12), generator=g).to(dev)
    grid = VoxelGrid(dens, feat, VoxelSize(0.5, 0.5, 0.5), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.Softplus(), expected_density_scale=2.0, tunable=True)
    out = grid(torch.zeros((0, 3), device=dev))
    assert out.shape == (0, 13)
    out.sum().backward()  # an empty query contributes zero gradients of the right shape
    assert float(grid.features.grad.abs().max()) == 0.0 and grid.densities.grad.shape == dens.shape
    one = grid(torch.tensor([[0.1, -0.2, 0.3]], device=dev))
    want = query_points_oracle(dens.cpu(), feat.cpu(), OracleGrid((0.5, 0.5, 0.5), density_scale=2.0, preact="identity", postact="softplus"),
                               torch.tensor([[0.1, -0.2, 0.3]]))
    assert one.shape == (1, 13) and float((one.detach().cpu().double() - want).abs().max()) <= 1e-5
    # the voxel centres themselves: the interpolation returns the stored values
    ax = [(torch.arange(n) + 0.5) * 0.5 - n * 0.25 for n in (3, 4, 5)]
    centres = torch.stack(torch.meshgrid(*ax, indexing="ij"), dim=-1).reshape(-1, 3).to(dev)
    at = grid(centres).detach()
    assert float((at[:, :12] - feat.reshape(-1, 12)).abs().max()) <= 1e-5
    assert float((at[:, 12] - torch.nn.functional.softplus(dens.reshape(-1) * 2.0)).abs().max()) <= 1e-5
    # frozen densities: only the features receive a gradient; NaN coordinates propagate instead of reading out of bounds
    frozen = VoxelGrid(dens.clone(), torch.nn.Parameter(feat.clone()), VoxelSize(0.5, 0.5, 0.5), tunable=False)
    res = frozen(torch.tensor([[0.0, 0.0, 0.0], [float("nan"), 0.0, 0.0], [1e30, -1e30, 0.0]], device=dev))
    assert torch.isfinite(res[0]).all() and torch.isnan(res[1]).any() and float(res[2, :12].abs().max()) == 0.0
    res[0].sum().backward()
    assert frozen.features.grad is not None and float(frozen.features.grad.abs().max()) > 0.0
    with pytest.raises(NotImplementedError, match="detach the points"):
        frozen(torch.zeros((2, 3), device=dev, requires_grad=True))
    with pytest.raises(AssertionError, match="attention grid"):
        frozen.forward_attn(torch.zeros((2, 3), device=dev))
