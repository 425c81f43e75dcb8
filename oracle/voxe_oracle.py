"""CPU oracle for the Vox-E ray-marching hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``vox-e_b200/``) never does, and fails loudly when the CUDA library is missing.

It is a from-scratch restatement (explicit gathers, no ``grid_sample``; closed loops over the eight
corners; one flat function) of what the reference computes for ``render_sh_voxel_grid``:

  * ray interval / AABB slab test ....... thre3d_atom/rendering/volumetric/sample.py:15-68, 71-184, 187-202
  * point normalisation .................. thre3d_atom/thre3d_reprs/voxels.py:225-234 and
                                           thre3d_atom/utils/imaging_utils.py:42-71 (``slack=True`` branch)
  * trilinear fetch (``grid_sample``,      thre3d_atom/thre3d_reprs/voxels.py:287-342 (PyTorch semantics:
    bilinear, zeros padding,               align_corners=False, out-of-range corners contribute zero)
    align_corners=False)
  * SH colour, inside-mask ............... thre3d_atom/rendering/volumetric/process.py:20-96,
                                           thre3d_atom/rendering/volumetric/utils/spherical_harmonics.py:64-132,
                                           thre3d_atom/thre3d_reprs/voxels.py:263-285
  * alpha compositing .................... thre3d_atom/rendering/volumetric/accumulate.py:24-113
  * ray casting (for harnesses) .......... thre3d_atom/rendering/volumetric/utils/misc.py:12-50,
                                           thre3d_atom/utils/imaging_utils.py:188-194

Parity pin: the reference's own tests hold no golden vectors for this path, so the oracle
is pinned against outputs of the *executed* reference: ``tests/golden/make_golden.py`` imports
``/root/reference`` in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks this file against every one of them (forward outputs and voxel
gradients).  Gradients come from torch autograd over this restatement (``dtype=torch.float64`` is the
"truth" mode, ``torch.float32`` mimics the reference's rounding).  It is written with whole-tensor torch ops like the
reference itself and runs on whatever device its inputs live on: the CPU everywhere it is used as a baseline, and -- for
the 15 GB grid of BASELINE.json's largest configuration only -- a CUDA device, where it is still plain ATen arithmetic
independent of the kernels under test.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

ZERO_PLUS = 1e-10  # thre3d_atom/utils/constants.py:8
INFINITY = 1e10  # thre3d_atom/utils/constants.py:9

# real SH constants, PlenOctrees convention (spherical_harmonics.py:33-50)
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (
    -0.5900435899266435,
    2.890611442640554,
    -0.4570457994644658,
    0.3731763325901154,
    -0.4570457994644658,
    1.445305721320277,
    -0.5900435899266435,
)


@dataclasses.dataclass
class OracleGrid:
    """Geometry + activation description of a voxel grid (voxels.py:46-130)."""

    voxel_size: Tuple[float, float, float]
    location: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    density_scale: float = 1.0
    preact: str = "abs"  # 'identity' | 'abs'          (VoxelGrid default is torch.abs, voxels.py:56)
    postact: str = "identity"  # 'identity' | 'relu' | 'softplus'


@dataclasses.dataclass
class OracleConfig:
    """The fields of SHVoxGridRenderConfig (renderers.py:29-47) that influence the arithmetic."""

    num_samples: int
    near: float
    far: float
    perturb: bool = False
    optimized_sampling: bool = False
    linear_disparity_sampling: bool = False
    white_bkgd: bool = False
    render_diffuse: bool = False
    noise_std: float = 0.0
    attn_mode: bool = False  # render_sh_voxel_grid_attn: 1 colour channel, background term forced to zero


def aabb_of(dims: Tuple[int, int, int], grid: OracleGrid) -> Tuple[Tuple[float, float], ...]:
    """voxels.py:198-223 -- python doubles, exactly as the reference builds them."""
    out = []
    for n, size, c in zip(dims, grid.voxel_size, grid.location):
        half = (n * size) / 2
        out.append((c - half, c + half))
    return tuple(out)


def sh_basis(degree: int, v: Tensor) -> Tensor:
    """Y_k(v) for k < (degree+1)^2 with the reference's signs (spherical_harmonics.py:86-116). v: [R,3] unit."""
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    cols = [torch.full_like(x, SH_C0)]
    if degree > 0:
        cols += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        cols += [
            SH_C2[0] * xy,
            SH_C2[1] * yz,
            SH_C2[2] * (2.0 * zz - xx - yy),
            SH_C2[3] * xz,
            SH_C2[4] * (xx - yy),
        ]
        if degree > 2:
            cols += [
                SH_C3[0] * y * (3 * xx - yy),
                SH_C3[1] * xy * z,
                SH_C3[2] * y * (4 * zz - xx - yy),
                SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy),
                SH_C3[4] * x * (4 * zz - xx - yy),
                SH_C3[5] * z * (xx - yy),
                SH_C3[6] * x * (xx - 3 * yy),
            ]
    return torch.stack(cols, dim=-1)


def ray_intervals(
    rays_o: Tensor, rays_d: Tensor, cfg: OracleConfig, aabb, dtype
) -> Tuple[Tensor, Tensor]:
    """Per-ray (near, far).  sample.py:38-44 for the plain case; sample.py:71-184 for the slab test."""
    R = rays_o.shape[0]
    dev = rays_o.device
    near = torch.full((R,), cfg.near, dtype=dtype, device=dev)
    far = torch.full((R,), cfg.far, dtype=dtype, device=dev)
    if not cfg.optimized_sampling:
        return near, far
    lo = hi = None
    hit = torch.ones(R, dtype=torch.bool, device=dev)
    for axis in range(3):
        denom = rays_d[:, axis] + ZERO_PLUS
        t0 = (aabb[axis][0] - rays_o[:, axis]) / denom
        t1 = (aabb[axis][1] - rays_o[:, axis]) / denom
        a_lo = torch.where(t0 > t1, t1, t0)
        a_hi = torch.where(t0 > t1, t0, t1)
        if axis == 0:
            lo, hi = a_lo, a_hi
            continue
        hit = hit & ~((lo > a_hi) | (a_lo > hi))
        lo = torch.where(a_lo > lo, a_lo, lo)
        hi = torch.where(a_hi < hi, a_hi, hi)
    lo = torch.where(hit, lo, near)
    hi = torch.where(hit, hi, far)
    return torch.clamp(lo, min=0.0), torch.clamp(hi, min=0.0)


def sample_depths(near: Tensor, far: Tensor, cfg: OracleConfig, jitter: Optional[Tensor], dtype) -> Tensor:
    """z_vals [R,S].  sample.py:46-64.  Disparity sampling is only reachable without optimized_sampling
    (renderers.py:66-78)."""
    S = cfg.num_samples
    t = torch.linspace(0.0, 1.0, S, dtype=dtype, device=near.device)[None, :]
    n, f = near[:, None], far[:, None]
    if cfg.linear_disparity_sampling and not cfg.optimized_sampling:
        z = 1.0 / (1.0 / (n + ZERO_PLUS) * (1.0 - t) + 1.0 / f * t)
    else:
        z = n * (1.0 - t) + f * t
    if cfg.perturb:
        assert jitter is not None, "perturb=True needs the stratified jitter u[R,S] (torch.rand in the reference)"
        mid = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mid, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mid], dim=-1)
        z = lower + (upper - lower) * jitter.to(dtype)
    return z


def _activate(x: Tensor, kind: str) -> Tensor:
    if kind == "identity":
        return x
    if kind == "abs":
        return torch.abs(x)
    if kind == "relu":
        return torch.relu(x)
    if kind == "softplus":  # torch.nn.Softplus(beta=1, threshold=20)
        return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))
    raise ValueError(kind)


def trilinear_fetch(vol: Tensor, pts: Tensor, aabb, dtype) -> Tensor:
    """vol [X,Y,Z,C], pts [N,3] world -> [N,C].

    Follows voxels.py:225-234 (normalise with numpy-fp32 scale/bias, imaging_utils.py:57-63) and then the
    published semantics of torch.nn.functional.grid_sample(mode='bilinear', padding_mode='zeros',
    align_corners=False): u = ((n + 1) * N - 1) / 2, corners floor(u), floor(u)+1, zero outside [0, N-1].
    The permute in voxels.py:308-311 makes point (x,y,z) index vol[ix,iy,iz].
    """
    dims = vol.shape[:3]
    C = vol.shape[3]
    idx0, frac = [], []
    for a in range(3):
        lo32, hi32 = np.float32(aabb[a][0]), np.float32(aabb[a][1])
        scale = (np.float32(1.0) - np.float32(-1.0)) / (hi32 - lo32)
        bias = np.float32(-1.0) - lo32 * scale
        n = pts[:, a] * float(scale) + float(bias)  # the fp32-rounded constants are part of the spec in both modes
        u = ((n + 1.0) * dims[a] - 1.0) / 2.0
        i0 = torch.floor(u)
        idx0.append(i0.long())
        frac.append(u - i0)
    out = torch.zeros(pts.shape[0], C, dtype=dtype, device=pts.device)
    flat = vol.reshape(-1, C)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                ix, iy, iz = idx0[0] + dx, idx0[1] + dy, idx0[2] + dz
                w = (
                    (frac[0] if dx else 1.0 - frac[0])
                    * (frac[1] if dy else 1.0 - frac[1])
                    * (frac[2] if dz else 1.0 - frac[2])
                )
                ok = (ix >= 0) & (ix < dims[0]) & (iy >= 0) & (iy < dims[1]) & (iz >= 0) & (iz < dims[2])
                lin = (ix.clamp(0, dims[0] - 1) * dims[1] + iy.clamp(0, dims[1] - 1)) * dims[2] + iz.clamp(
                    0, dims[2] - 1
                )
                out = out + torch.where(ok, w, torch.zeros_like(w))[:, None] * flat[lin]
    return out


def query_points_oracle(densities: Tensor, features: Tensor, grid: OracleGrid, points: Tensor, dtype=torch.float64) -> Tensor:
    """VoxelGrid.forward / forward_attn (voxels.py:287-345, 347-406): [N, F + 1] = (interpolated features, post(interpolated
    pre(densities * scale))) at world-space ``points`` [N,3] -- anywhere, zeros padding, no inside mask.  Pass the attention
    grid [X,Y,Z,1] as ``features`` for forward_attn.  Differentiable w.r.t. both grid tensors."""
    densities, features = densities.to(dtype), features.to(dtype)
    aabb = aabb_of(tuple(features.shape[:3]), grid)
    pts = points.to(torch.float32).to(dtype)
    pre = _activate(densities * grid.density_scale, grid.preact)  # on the voxels, then interpolated (voxels.py:303-305)
    sigma = _activate(trilinear_fetch(pre, pts, aabb, dtype), grid.postact)
    return torch.cat([trilinear_fetch(features, pts, aabb, dtype), sigma], dim=-1)


def render_oracle(
    densities: Tensor,
    features: Tensor,
    grid: OracleGrid,
    rays_o: Tensor,
    rays_d: Tensor,
    cfg: OracleConfig,
    jitter: Optional[Tensor] = None,
    noise: Optional[Tensor] = None,
    dtype: torch.dtype = torch.float64,
) -> Dict[str, Tensor]:
    """Whole path, differentiable w.r.t. ``densities`` [X,Y,Z,1] and ``features`` [X,Y,Z,F].

    Returns dict(colour [R,3 or 1], depth [R,1], disparity [R,1], accumulated_weight [R,1], inside [R,S]).
    """
    densities, features = densities.to(dtype), features.to(dtype)
    R, S = rays_o.shape[0], cfg.num_samples
    dims = tuple(features.shape[:3])
    aabb = aabb_of(dims, grid)
    n_col = 1 if cfg.attn_mode else 3

    # Geometry is DEFINED in fp32, op for op as the reference evaluates it (separate mul and add, IEEE divide):
    # with optimized_sampling the first/last sample sits exactly on an AABB face, so whether it counts as inside
    # (strict compare, step 7) -- and with it a delta of 1e10 -- is decided by fp32 rounding.  Only the
    # interpolation / SH / compositing arithmetic below runs in ``dtype``.
    ro32, rd32 = rays_o.to(torch.float32), rays_d.to(torch.float32)
    near, far = ray_intervals(ro32, rd32, cfg, aabb, torch.float32)
    z32 = sample_depths(near, far, cfg, jitter, torch.float32)  # [R,S]
    pts32 = (ro32[:, None, :] + rd32[:, None, :] * z32[:, :, None]).reshape(-1, 3)
    inside = torch.ones(pts32.shape[0], dtype=torch.bool, device=pts32.device)
    for a in range(3):
        inside &= (pts32[:, a] > aabb[a][0]) & (pts32[:, a] < aabb[a][1])  # python double vs fp32 tensor, as voxels.py:263-285
    inside = inside.reshape(R, S)
    z, pts = z32.to(dtype), pts32.to(dtype)
    rays_o, rays_d = ro32.to(dtype), rd32.to(dtype)

    # density: pre-activation on the *voxels*, interpolate, post-activation (voxels.py:303-320)
    pre = _activate(densities * grid.density_scale, grid.preact)
    sigma = _activate(trilinear_fetch(pre, pts, aabb, dtype), grid.postact)[:, 0]
    feats = trilinear_fetch(features, pts, aabb, dtype)  # [N,F]

    # SH colour (process.py:46-76); coefficient layout is channel-major f = c*K + k
    K = features.shape[-1] // n_col
    degree = int(round(math.sqrt(K))) - 1
    assert (degree + 1) ** 2 == K and 0 <= degree <= 3
    vdir = rays_d / torch.linalg.norm(rays_d, dim=-1, keepdim=True)
    coef = feats.reshape(R, S, n_col, K)
    if cfg.render_diffuse:
        raw = SH_C0 * coef[..., 0]
    else:
        raw = (coef * sh_basis(degree, vdir)[:, None, None, :]).sum(-1)  # [R,S,n_col]

    # strict inside test on the world-space points (voxels.py:263-285; process.py:80-84): mask computed above
    raw = torch.where(inside[..., None], raw, torch.full_like(raw, -INFINITY))
    sigma = torch.where(inside, sigma.reshape(R, S), torch.zeros(R, S, dtype=dtype, device=sigma.device))

    # compositing (accumulate.py:49-88)
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.full((R, 1), INFINITY, dtype=dtype, device=z.device)], dim=-1)
    delta = delta * torch.linalg.norm(rays_d, dim=-1, keepdim=True)
    if cfg.noise_std != 0.0:
        assert noise is not None
        sigma = sigma + noise.to(dtype) * cfg.noise_std
    alpha = 1.0 - torch.exp(-(sigma * delta))
    trans = torch.cumprod(torch.cat([torch.ones(R, 1, dtype=dtype, device=alpha.device), 1.0 - alpha], dim=-1), dim=-1)[:, :-1]
    w = alpha * trans
    colour = (torch.sigmoid(raw) * w[..., None]).sum(dim=1)
    acc = w.sum(dim=-1, keepdim=True)
    if cfg.white_bkgd and not cfg.attn_mode:
        colour = colour + (1.0 - acc)
    depth = (z * w).sum(dim=-1, keepdim=True)
    disparity = 1.0 / torch.maximum(torch.full_like(acc, ZERO_PLUS), depth / acc)
    return {
        "colour": colour,
        "depth": depth,
        "disparity": disparity,
        "accumulated_weight": acc,
        "inside": inside,
    }


def relu_kink_voxels(
    densities: Tensor,
    grid: OracleGrid,
    rays_o: Tensor,
    rays_d: Tensor,
    cfg: OracleConfig,
    jitter: Optional[Tensor] = None,
    margin: float = 2e-3,
) -> Tensor:
    """Bool mask [X,Y,Z] of the voxels that are trilinear corners of an in-grid sample whose interpolated
    (pre-activated) density lies within ``margin`` of 0.

    With a ReLU post-activation such a sample's derivative is decided by fp32 rounding (the reference, the fp64 truth
    and any re-ordered fp32 evaluation may disagree), and a flip changes only that sample's own scatter into its 8
    corner voxels of d_densities -- forward values and every other gradient entry move by < margin * delta.  Parity
    tests therefore compare d_densities on the complement of this mask at the tight tolerance."""
    dtype = torch.float64
    dims = tuple(densities.shape[:3])
    aabb = aabb_of(dims, grid)
    ro32, rd32 = rays_o.to(torch.float32), rays_d.to(torch.float32)
    near, far = ray_intervals(ro32, rd32, cfg, aabb, torch.float32)
    z32 = sample_depths(near, far, cfg, jitter, torch.float32)
    pts32 = (ro32[:, None, :] + rd32[:, None, :] * z32[:, :, None]).reshape(-1, 3)
    inside = torch.ones(pts32.shape[0], dtype=torch.bool, device=pts32.device)
    for a in range(3):
        inside &= (pts32[:, a] > aabb[a][0]) & (pts32[:, a] < aabb[a][1])
    pts = pts32.to(dtype)
    pre = _activate(densities.to(dtype) * grid.density_scale, grid.preact)
    sraw = trilinear_fetch(pre, pts, aabb, dtype)[:, 0]
    amb = inside & (sraw.abs() < margin)
    mask = torch.zeros(dims, dtype=torch.bool, device=pts32.device)
    if not amb.any():
        return mask
    p = pts[amb]
    idx0 = []
    for a in range(3):
        lo32, hi32 = np.float32(aabb[a][0]), np.float32(aabb[a][1])
        scale = (np.float32(1.0) - np.float32(-1.0)) / (hi32 - lo32)
        bias = np.float32(-1.0) - lo32 * scale
        u = ((p[:, a] * float(scale) + float(bias) + 1.0) * dims[a] - 1.0) / 2.0
        idx0.append(torch.floor(u).long())
    for dx in (0, 1):  # the 2x2x2 footprint (a corner reached only through u's rounding noise has weight ~1e-5)
        for dy in (0, 1):
            for dz in (0, 1):
                ix = (idx0[0] + dx).clamp(0, dims[0] - 1)
                iy = (idx0[1] + dy).clamp(0, dims[1] - 1)
                iz = (idx0[2] + dz).clamp(0, dims[2] - 1)
                mask[ix, iy, iz] = True
    return mask


def render_oracle_with_grads(
    densities: Tensor,
    features: Tensor,
    grid: OracleGrid,
    rays_o: Tensor,
    rays_d: Tensor,
    cfg: OracleConfig,
    g_colour: Tensor,
    g_depth: Optional[Tensor] = None,
    g_acc: Optional[Tensor] = None,
    g_disp: Optional[Tensor] = None,
    jitter: Optional[Tensor] = None,
    noise: Optional[Tensor] = None,
    dtype: torch.dtype = torch.float64,
) -> Dict[str, Tensor]:
    """Forward + autograd backward of L = <g_colour,colour> + <g_depth,depth> + <g_acc,acc> + <g_disp,disparity>."""
    d = densities.detach().clone().to(dtype).requires_grad_(True)
    f = features.detach().clone().to(dtype).requires_grad_(True)
    out = render_oracle(d, f, grid, rays_o, rays_d, cfg, jitter=jitter, noise=noise, dtype=dtype)
    loss = (out["colour"] * g_colour.to(dtype)).sum()
    if g_depth is not None:
        loss = loss + (out["depth"] * g_depth.to(dtype)).sum()
    if g_acc is not None:
        loss = loss + (out["accumulated_weight"] * g_acc.to(dtype)).sum()
    if g_disp is not None:
        loss = loss + (out["disparity"] * g_disp.to(dtype)).sum()
    loss.backward()
    res = {k: v.detach() for k, v in out.items()}
    res["d_densities"] = d.grad.detach()
    res["d_features"] = f.grad.detach()
    return res


# ---------------------------------------------------------------------------------------------------------
# harness helpers (camera model) -- restated from misc.py:12-50 and imaging_utils.py:153-194
# ---------------------------------------------------------------------------------------------------------
def pose_spherical_np(yaw_deg: float, pitch_deg: float, radius: float) -> Tuple[np.ndarray, np.ndarray]:
    """c2w = Rz(yaw) @ Rx(pitch) @ Tz(radius), built in fp32 like the reference. -> (rotation [3,3], translation [3,1])"""
    yaw, pitch = yaw_deg / 180.0 * np.pi, pitch_deg / 180.0 * np.pi
    tz = np.eye(4, dtype=np.float32)
    tz[2, 3] = radius
    rx = np.array(
        [[1, 0, 0, 0], [0, np.cos(pitch), -np.sin(pitch), 0], [0, np.sin(pitch), np.cos(pitch), 0], [0, 0, 0, 1]],
        dtype=np.float32,
    )
    rz = np.array(
        [[np.cos(yaw), -np.sin(yaw), 0, 0], [np.sin(yaw), np.cos(yaw), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
        dtype=np.float32,
    )
    c2w = rz @ (rx @ tz)
    return c2w[:3, :3].copy(), c2w[:3, 3:].copy()


def cast_rays_np(height: int, width: int, focal: float, rotation: np.ndarray, translation: np.ndarray):
    """Pixel-centre pinhole rays, flat index = y*W + x, directions NOT normalised. -> (origins, directions) [H*W,3] fp32"""
    xs = torch.linspace(0.5, width - 0.5, width, dtype=torch.float32)
    ys = torch.linspace(0.5, height - 0.5, height, dtype=torch.float32)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    dirs = torch.stack([(xx - width * 0.5) / focal, -(yy - height * 0.5) / focal, -torch.ones_like(xx)], dim=-1)
    rot = torch.as_tensor(rotation, dtype=torch.float32)
    d = (rot @ dirs[..., None])[..., 0].reshape(-1, 3)
    o = torch.as_tensor(translation, dtype=torch.float32).reshape(1, 3).expand_as(d).contiguous()
    return o, d.contiguous()
