"""Row f2, regularisers: the total-variation and density-correlation / L2 / L1 losses of the edit loop
(thre3d_atom/modules/sds_trainer.py:290-326, 494-524, 563-567).

CPU: the oracle restatement (``oracle/regularizers_oracle.py``) against golden vectors produced by executing the reference's
own function bodies (``tests/golden/make_golden_regularizers.py``).  GPU: the streaming kernels, through the
reference-named functions of ``voxe_b200.regularizers`` (-> C ABI), against the same goldens, against the fp64 oracle on
seeded grids, and at the headline grid size through size-independent properties.

Tolerances (fp32 path): loss |d| <= 2e-6 + 1e-5 * |loss|; gradient max-abs <= 1e-5 * ||g||_inf (the TV gradient is a sum of
signs times three constants, exact up to the rounding of those constants).
"""
import json

import numpy as np
import pytest
import torch

from _golden import GOLDEN_DIR
from oracle import regularizers_oracle as orc


def _load():
    z = np.load(GOLDEN_DIR / "regularizers.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}


META, G = _load()


def _close_loss(got, want, what):
    got, want = float(got.detach()), float(want)
    assert abs(got - want) <= 2e-6 + 1e-5 * abs(want), (what, got, want)


def _close_grad(got, want, what, tol=1e-5):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs().max().item()
    assert err <= tol * max(want.abs().max().item(), 1e-30), (what, err, want.abs().max().item())


# ---------------------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", META["tv"], ids=lambda c: c["name"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_oracle_tv_matches_reference(case, dtype):
    name = case["name"]
    got = orc.with_grad(orc.tv_loss, G[f"tv_{name}_grid"], relu=case["relu"], upstream=case["upstream"], dtype=dtype)
    _close_loss(got["loss"], G[f"tv_{name}_loss"], name)
    _close_grad(got["grad"], G[f"tv_{name}_grad"], name)


@pytest.mark.parametrize("case", META["pair"], ids=lambda c: c["name"])
def test_oracle_pair_loss_matches_reference(case):
    name, mode = case["name"], case["mode"]
    a, b = G[f"pair_{name}_a"], G[f"pair_{name}_b"]
    got = orc.with_grad(orc.density_pair_loss, a, b, mode=mode, upstream=case["upstream"], dtype=torch.float64)
    _close_loss(got["loss"], G[f"pair_{name}_loss"], name)
    # the reference's fp32 autograd of 1 - cov/(sqrt(var var) + eps) near correlation 1 cancels two nearly equal terms
    _close_grad(got["grad"], G[f"pair_{name}_grad"], name, tol=2e-4 if mode == "correlation" else 1e-5)
    if mode == "correlation":
        _close_grad(orc.correlation_grid(a.double(), b.double()), G[f"pair_{name}_corr"], name)


def test_product_refuses_cpu_tensors():
    from voxe_b200 import regularizers as reg

    x = torch.zeros(4, 4, 4, 1, requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA only"):
        reg._tv_loss_on_grid(x)
    with pytest.raises(RuntimeError, match="CUDA only"):
        reg.density_correlation_loss_fn(x, torch.zeros(4, 4, 4, 1))
    with pytest.raises(ValueError):
        reg._tv_loss_on_grid(torch.zeros(4, 4, 4))


def test_argument_validation_without_a_gpu():
    import ctypes

    from voxe_b200 import _native as nat

    lib = nat.load_library()
    dims = (ctypes.c_int32 * 3)(4, 0, 4)
    assert lib.voxe_tv_regularizer(1, ctypes.byref(dims), 1, 0, None, None, None, 1.0, None, 0, None) == 1
    assert b"dims" in lib.voxe_last_error()
    dims = (ctypes.c_int32 * 3)(4, 4, 4)
    assert lib.voxe_tv_regularizer(1, ctypes.byref(dims), 1, 0, None, None, None, 1.0, None, 0, None) == 0  # nothing asked for
    assert lib.voxe_tv_regularizer(1, ctypes.byref(dims), 1, 0, None, 1, None, 1.0, None, 0, None) == 1
    assert b"workspace" in lib.voxe_last_error()
    assert lib.voxe_pair_loss(None, None, 8, 0, None, None, None, None) == 1
    assert lib.voxe_pair_loss(1, 1, 8, 7, 1, 1, None, None) == 1 and b"mode" in lib.voxe_last_error()
    assert lib.voxe_pair_loss_grad(1, 1, 8, 0, None, None, 1.0, 1, 0, None) == 1 and b"workspace" in lib.voxe_last_error()


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", META["tv"], ids=lambda c: c["name"])
def test_tv_kernel_matches_reference_golden(case):
    from voxe_b200 import _native as nat
    from voxe_b200 import regularizers as reg

    name = case["name"]
    x = G[f"tv_{name}_grid"].cuda().requires_grad_(True)
    before = nat.launch_count()
    loss = reg._tv_loss_on_grid(x, relu=case["relu"])
    (loss * case["upstream"]).backward()
    assert nat.launch_count() - before == 2  # loss pass (its last CTA folds the partial sums), gradient pass
    _close_loss(loss, G[f"tv_{name}_loss"], name)
    _close_grad(x.grad, G[f"tv_{name}_grad"], name)
    # the no-autograd route: one pass, accumulates on top of what .grad holds
    y = G[f"tv_{name}_grid"].cuda().requires_grad_(True)
    base = torch.randn_like(y) * G[f"tv_{name}_grad"].abs().max().item()
    y.grad = base.clone()
    loss2 = reg.accumulate_tv_gradient(y, case["upstream"], relu=case["relu"])
    _close_loss(loss2, G[f"tv_{name}_loss"], name)
    _close_grad(y.grad - base, G[f"tv_{name}_grad"], name, tol=1e-5 + 4 * 6e-8)  # the sum is rounded once more


@pytest.mark.gpu
@pytest.mark.parametrize("case", META["pair"], ids=lambda c: c["name"])
def test_pair_kernels_match_reference_golden(case):
    from voxe_b200 import regularizers as reg

    name, mode = case["name"], case["mode"]
    a = G[f"pair_{name}_a"].cuda().requires_grad_(True)
    b = G[f"pair_{name}_b"].cuda()
    loss, corr = reg.density_correlation_loss_fn(sds_density=a, regular_density=b, l2_mode=mode == "l2", l1_mode=mode == "l1")
    (loss * case["upstream"]).backward()
    _close_loss(loss, G[f"pair_{name}_loss"], name)
    want = orc.with_grad(orc.density_pair_loss, G[f"pair_{name}_a"], G[f"pair_{name}_b"], mode=mode, upstream=case["upstream"])
    _close_grad(a.grad, want["grad"], name)  # against fp64 truth
    _close_grad(a.grad, G[f"pair_{name}_grad"], name, tol=2e-4 if mode == "correlation" else 1e-5)  # against the fp32 reference
    if mode == "correlation":
        assert corr.shape == a.shape and not corr.requires_grad
        _close_grad(corr, G[f"pair_{name}_corr"], name)
    else:
        assert corr is None
    a2 = G[f"pair_{name}_a"].cuda().requires_grad_(True)
    loss2 = reg.accumulate_density_loss_gradient(a2, b, case["upstream"], mode=mode)
    _close_loss(loss2, G[f"pair_{name}_loss"], name)
    _close_grad(a2.grad, want["grad"], name)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,channels,relu", [((33, 17, 40), 1, True), ((16, 16, 16), 3, False), ((5, 9, 300), 12, False),
                                                ((3, 1, 7), 2, False), ((1, 6, 6), 1, True), ((20, 20, 1), 4, False),
                                                ((6, 5, 8), 27, False), ((4, 7, 4), 5, True), ((9, 4, 6), 2, True),
                                                ((7, 3, 5), 3, False), ((300, 2, 2), 48, False)])
def test_tv_kernel_against_fp64_oracle(dims, channels, relu):
    """Seeded grids; every kernel variant (16-byte path with the z neighbours in the window C = 1, 2, 3, as aligned groups
    C % 4 == 0, as scalars C = 5 / 27; scalar path when a row is not a multiple of 16 bytes), row lengths below / above one
    CTA stride, and axes of extent 1 (loss NaN as torch's empty mean, finite gradient from the other axes)."""
    from voxe_b200 import regularizers as reg

    g = torch.Generator().manual_seed(sum(dims) + channels)
    x = torch.randn((*dims, channels), generator=g)
    x[::2, ::3] = torch.round(x[::2, ::3])  # ties
    xc = x.cuda().requires_grad_(True)
    loss = reg._tv_loss_on_grid(xc, relu=relu)
    loss.backward()
    if 1 in dims:
        assert torch.isnan(loss)
        xd = x.double().requires_grad_(True)
        h = torch.relu(xd) if relu else xd
        parts = [h.diff(dim=a).abs().mean() for a in range(3) if dims[a] > 1]
        (sum(parts) / 3).backward()
        _close_grad(xc.grad, xd.grad, str(dims))
        return
    want = orc.with_grad(orc.tv_loss, x, relu=relu)
    _close_loss(loss, want["loss"], str(dims))
    _close_grad(xc.grad, want["grad"], str(dims))


@pytest.mark.gpu
def test_headline_grid_properties():
    """160^3 (BASELINE.json config 2/3 grid): properties that need no CPU pass over 4 M voxels."""
    from voxe_b200 import regularizers as reg

    gen = torch.Generator(device="cuda").manual_seed(3)
    dens = torch.randn((160, 160, 160, 1), device="cuda", generator=gen).requires_grad_(True)
    feat = torch.randn((160, 160, 160, 3), device="cuda", generator=gen).requires_grad_(True)
    # TV is translation invariant and positively homogeneous; its gradient is scale invariant and sums to zero
    for x in (dens, feat):
        l1 = reg._tv_loss_on_grid(x)
        l1.backward()
        g1 = x.grad.clone()
        x.grad = None
        l2 = reg._tv_loss_on_grid((x.detach() * 4.0 + 1.5).requires_grad_(True))
        assert abs(float(l2.detach()) - 4.0 * float(l1.detach())) <= 1e-5 * float(l2.detach())
        assert abs(float(g1.double().sum())) <= 1e-6
        # E|N(0,1) - N(0,1)| = 2 / sqrt(pi)
        assert abs(float(l1.detach()) - 2.0 / np.pi ** 0.5) < 2e-3
        ref = x.detach().clone().requires_grad_(True)  # torch's own ops on the GPU: same formula, independent kernels
        ((ref.diff(dim=0).abs().mean() + ref.diff(dim=1).abs().mean() + ref.diff(dim=2).abs().mean()) / 3).backward()
        _close_grad(g1, ref.grad, "tv 160^3")
    # correlation loss: 0 against itself (up to eps), 2 against its negation, invariant to affine maps of either grid
    base = dens.detach()
    l_self, _ = reg.density_correlation_loss_fn(dens, base)
    l_neg, _ = reg.density_correlation_loss_fn(dens, -base)
    l_aff, _ = reg.density_correlation_loss_fn(dens, base * 7.0 - 3.0, return_correlation_grid=False)
    assert abs(float(l_self)) < 1e-5 and abs(float(l_neg) - 2.0) < 1e-5 and abs(float(l_aff)) < 1e-5
    other = torch.randn(dens.shape, device="cuda", generator=gen)
    mixed = (0.6 * base + 0.8 * other).requires_grad_(True)
    loss, corr = reg.density_correlation_loss_fn(mixed, base)
    loss.backward()
    assert abs(float(loss) - 0.4) < 5e-3  # correlation 0.6
    assert abs(float(corr.double().mean()) - (1.0 - float(loss))) < 1e-5  # the loss is 1 - mean(correlation_grid)
    assert abs(float(mixed.grad.double().sum())) < 1e-6  # invariance to a constant shift
    assert abs(float((mixed.grad.double() * mixed.detach().double()).sum())) < 1e-4  # ... and to a rescale (Euler)
    ref = mixed.detach().clone().requires_grad_(True)
    da, db = ref - ref.mean(), base - base.mean()
    (1.0 - (da * db / (torch.sqrt((da ** 2).mean() * (db ** 2).mean()) + 1e-7)).mean()).backward()
    _close_grad(mixed.grad, ref.grad, "corr 160^3", tol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("dims,channels", [((256, 256, 256), 4), ((512, 512, 512), 1)])
def test_large_grids_against_torch_ops(dims, channels):
    """BASELINE.json's larger grids (256^3 with 4 channels = the packed SH-0 voxel; 512^3 densities, 537 MB): 64-bit indexing
    and the persistent spans at full size, checked against the same formula in torch ops on the GPU."""
    from voxe_b200 import regularizers as reg

    gen = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn((*dims, channels), device="cuda", generator=gen)
    x[:, ::5] = torch.round(x[:, ::5])
    ours = x.clone().requires_grad_(True)
    loss = reg._tv_loss_on_grid(ours, relu=True)
    loss.backward()
    ref = x.clone().requires_grad_(True)
    h = torch.relu(ref)
    want = (h.diff(dim=0).abs().mean() + h.diff(dim=1).abs().mean() + h.diff(dim=2).abs().mean()) / 3
    want.backward()
    _close_loss(loss, want.detach(), str(dims))
    assert (ours.grad - ref.grad).abs().max().item() <= 1e-5 * ref.grad.abs().max().item()
    del h, want
    other = (0.5 * x + torch.randn_like(x)).requires_grad_(True)
    l2, _ = reg.density_correlation_loss_fn(other, x, return_correlation_grid=False)
    l2.backward()
    o = other.detach().clone().requires_grad_(True)
    da, db = o - o.mean(), x - x.mean()
    w2 = 1.0 - (da * db / (torch.sqrt((da ** 2).mean() * (db ** 2).mean()) + 1e-7)).mean()
    w2.backward()
    _close_loss(l2, w2.detach(), str(dims))
    assert (other.grad - o.grad).abs().max().item() <= 1e-4 * o.grad.abs().max().item()


@pytest.mark.gpu
def test_regulariser_gradients_reach_the_fused_optimiser_step():
    """``accumulate_*`` leave dense gradients in ``.grad``; FusedVoxelAdam consumes them beside the render's packed volume
    exactly as torch.optim.Adam consumes autograd's."""
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from voxe_b200 import regularizers as reg
    from voxe_b200.optim import FusedVoxelAdam

    g = torch.Generator().manual_seed(5)
    dims = (12, 10, 14)
    dens, feat = torch.randn((*dims, 1), generator=g), torch.randn((*dims, 3), generator=g)
    pre = dens + 0.3 * torch.randn((*dims, 1), generator=g)
    grid = VoxelGrid(dens.clone().cuda(), feat.clone().cuda(), VoxelSize(*(3.0 / d for d in dims)), tunable=True)
    ours = FusedVoxelAdam(grid, lr=0.02)
    rd, rf = torch.nn.Parameter(dens.clone().cuda()), torch.nn.Parameter(feat.clone().cuda())
    ref = torch.optim.Adam([rd, rf], lr=0.02)
    for _ in range(3):
        reg.accumulate_density_loss_gradient(grid.densities, pre.cuda(), 200.0)
        reg.accumulate_tv_gradient(grid.densities, 0.5, relu=True)
        reg.accumulate_tv_gradient(grid.features, 0.1)
        ours.step()
        ours.zero_grad()
        total = 200.0 * orc.density_pair_loss(rd, pre.cuda()) + 0.5 * orc.tv_loss(rd, relu=True) + 0.1 * orc.tv_loss(rf)
        ref.zero_grad()
        total.backward()
        ref.step()
    assert torch.allclose(grid.densities, rd, atol=2e-4), float((grid.densities - rd).abs().max())
    assert torch.allclose(grid.features, rf, atol=2e-4), float((grid.features - rf).abs().max())
