"""Generate golden vectors by EXECUTING the unmodified reference (``/root/reference``) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference imports ``matplotlib`` (imaging_utils.py:4) and ``easydict`` (utils/misc.py:6), neither of which is
installed here and neither of which touches the render path; empty stub modules stand in for them.  Everything
else -- ``VoxelGrid``, ``render_sh_voxel_grid``, ``VolumetricModel``, ``cast_rays``, ``pose_spherical`` -- is the
reference's own code, fp32, PyTorch CPU.

Each ``render_*.npz`` case stores the inputs (grid tensors, geometry, rays, config, the stratified jitter the
reference drew, upstream gradients) and the reference's outputs (colour, depth, disparity, accumulated_weight,
d_densities, d_features).  ``cameras.npz`` pins ``pose_spherical`` / ``cast_rays``; ``volmodel.npz`` pins
``VolumetricModel.render`` (chunk loop + collation).  The jitter is captured by replaying the generator:
``sample.py:63`` is the first RNG draw of a render call, so ``torch.manual_seed(s); torch.rand(R, S)`` reproduces it.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
OUT_DIR = Path(__file__).resolve().parent


def _import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "easydict"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["easydict"].EasyDict = dict
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REFERENCE_ROOT)
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid, render_sh_voxel_grid_attn
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelGridLocation, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, get_thre360_animation_poses, pose_spherical

    return dict(locals())


ACT = {
    "identity": lambda: torch.nn.Identity(),
    "abs": lambda: torch.abs,
    "relu": lambda: torch.nn.ReLU(),
    "softplus": lambda: torch.nn.Softplus(),
}


def blob_grid(dims, n_feat, rng):
    """Smooth semi-transparent content: density falls off with radius, features are low-frequency sinusoids
    plus a little noise so every channel is distinct."""
    X, Y, Z = dims
    gx, gy, gz = np.meshgrid(
        np.linspace(-1, 1, X), np.linspace(-1, 1, Y), np.linspace(-1, 1, Z), indexing="ij"
    )
    r = np.sqrt(gx**2 + gy**2 + gz**2)
    dens = (0.6 * (1.0 - r) + 0.05 * rng.standard_normal(r.shape))[..., None]
    ch = []
    for k in range(n_feat):
        ch.append(np.sin((1 + k % 3) * gx + 0.3 * k) * np.cos((1 + k % 2) * gy - 0.2 * k) + 0.5 * gz + 0.1 * rng.standard_normal(r.shape))
    feat = np.stack(ch, axis=-1)
    return dens.astype(np.float32), feat.astype(np.float32)


def uniform_grid(dims, n_feat, rng):
    dens = rng.uniform(-1, 1, (*dims, 1)).astype(np.float32)
    feat = rng.uniform(-1, 1, (*dims, n_feat)).astype(np.float32)
    return dens, feat


def camera_rays(ref, height, width, focal, yaw, pitch, radius):
    pose = ref["pose_spherical"](yaw, pitch, radius)
    rays = ref["flatten_rays"](ref["cast_rays"](ref["CameraIntrinsics"](height, width, focal), pose))
    return rays.origins.contiguous().clone(), rays.directions.contiguous().clone()


def special_rays(aabb_lo, aabb_hi):
    """Hand-built probes: misses, origin inside the grid, axis-parallel directions (d_a == 0), grazing a face."""
    c = 0.5 * (np.asarray(aabb_lo) + np.asarray(aabb_hi))
    e = 0.5 * (np.asarray(aabb_hi) - np.asarray(aabb_lo))
    o, d = [], []
    o.append(c + [0, 0, 4 * e[2]]); d.append([0.0, 0.0, -1.0])  # straight down the z axis (dx=dy=0)
    o.append(c + [4 * e[0], 0.1 * e[1], 0.2 * e[2]]); d.append([-1.0, 0.0, 0.0])  # along -x
    o.append(c + [0.3 * e[0], -4 * e[1], 0.0]); d.append([0.0, 1.0, 0.0])  # along +y
    o.append(c + [0.1 * e[0], 0.2 * e[1], -0.3 * e[2]]); d.append([0.3, -0.5, 0.8])  # starts inside
    o.append(c + [0.0, 0.0, 0.0]); d.append([-1.0, 0.2, 0.1])  # starts at the centre
    o.append(c + [5 * e[0], 5 * e[1], 5 * e[2]]); d.append([1.0, 1.0, 1.0])  # points away: miss
    o.append(c + [0, 3 * e[1], 4 * e[2]]); d.append([0.0, 0.0, -1.0])  # parallel, outside the slab: miss
    o.append(c + [0.999 * e[0], 0.0, 4 * e[2]]); d.append([0.0, 0.0, -1.5])  # grazes the +x face, |d| != 1
    o.append(c + [3 * e[0], 3 * e[1], 3 * e[2]]); d.append([-1.0, -1.0, -1.0])  # through the diagonal
    o.append(c + [-4 * e[0], 0.5 * e[1], -0.5 * e[2]]); d.append([2.0, -0.1, 0.1])  # long direction vector
    return torch.tensor(np.asarray(o), dtype=torch.float32), torch.tensor(np.asarray(d), dtype=torch.float32)


# name, dims, sh_degree, voxel_size, location, grid kind, density_scale, pre, post, rays, S, cfg overrides, grads
CASES = [
    dict(name="relu_white", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="uniform", scale=33.333,
         pre="identity", post="relu", cam=(12, 12, 16.0, 30.0, 60.0, 4.0311), S=64, near=1.8, far=6.6, white=True),
    dict(name="softplus_white", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="uniform", scale=33.333,
         pre="identity", post="softplus", cam=(12, 12, 16.0, 100.0, 40.0, 4.0311), S=64, near=1.8, far=6.6, white=True),
    dict(name="abs_identity_black", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="blob", scale=1.0,
         pre="abs", post="identity", cam=(12, 12, 16.0, 200.0, 75.0, 4.0311), S=64, near=1.8, far=6.6, white=False),
    dict(name="identity_identity_far_outside", dims=(12, 12, 12), deg=0, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="blob", scale=2.0,
         pre="identity", post="identity", cam=(10, 10, 13.0, 10.0, 50.0, 4.0311), S=48, near=1.8, far=6.6, white=True),
    dict(name="abs_relu", dims=(12, 12, 12), deg=0, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="uniform", scale=5.0,
         pre="abs", post="relu", cam=(10, 10, 13.0, 300.0, 20.0, 4.0311), S=48, near=1.8, far=6.6, white=True),
    dict(name="abs_softplus", dims=(12, 12, 12), deg=0, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="blob", scale=3.0,
         pre="abs", post="softplus", cam=(10, 10, 13.0, 150.0, 65.0, 4.0311), S=48, near=1.8, far=6.6, white=False),
    dict(name="blob_relu_all_grads", dims=(20, 20, 20), deg=0, vsize=(0.15,) * 3, loc=(0, 0, 0), kind="blob", scale=8.0,
         pre="identity", post="relu", cam=(9, 11, 12.0, 45.0, 55.0, 4.0311), S=96, near=1.8, far=6.6, white=True,
         all_grads=True),
    dict(name="blob_softplus_all_grads_black", dims=(20, 20, 20), deg=0, vsize=(0.15,) * 3, loc=(0, 0, 0), kind="blob", scale=8.0,
         pre="identity", post="softplus", cam=(9, 11, 12.0, 250.0, 35.0, 4.0311), S=96, near=1.8, far=6.6, white=False,
         all_grads=True),
    dict(name="aabb_sampling", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="blob", scale=10.0,
         pre="identity", post="relu", cam=(12, 12, 14.0, 60.0, 60.0, 4.0311), S=64, near=1.8, far=6.6, white=True,
         optimized=True),
    dict(name="aabb_sampling_special_rays", dims=(10, 12, 14), deg=0, vsize=(0.3, 0.25, 0.2), loc=(0.2, -0.1, 0.3), kind="blob",
         scale=6.0, pre="identity", post="softplus", cam=None, S=32, near=0.5, far=9.0, white=True, optimized=True),
    dict(name="special_rays_plain", dims=(10, 12, 14), deg=0, vsize=(0.3, 0.25, 0.2), loc=(0.2, -0.1, 0.3), kind="blob",
         scale=6.0, pre="identity", post="relu", cam=None, S=40, near=0.5, far=9.0, white=True),
    dict(name="disparity_sampling", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="blob", scale=10.0,
         pre="identity", post="relu", cam=(12, 12, 16.0, 120.0, 60.0, 4.0311), S=64, near=1.8, far=6.6, white=True,
         disparity=True),
    dict(name="perturb_jitter", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="blob", scale=10.0,
         pre="identity", post="relu", cam=(12, 12, 16.0, 80.0, 60.0, 4.0311), S=64, near=1.8, far=6.6, white=True,
         perturb=True),
    dict(name="perturb_jitter_aabb", dims=(16, 16, 16), deg=0, vsize=(3 / 16,) * 3, loc=(0, 0, 0), kind="blob", scale=10.0,
         pre="identity", post="softplus", cam=(12, 12, 16.0, 170.0, 30.0, 4.0311), S=64, near=1.8, far=6.6, white=True,
         perturb=True, optimized=True),
    dict(name="anisotropic_offcentre", dims=(9, 14, 11), deg=0, vsize=(0.31, 0.2, 0.27), loc=(0.3, -0.2, 0.1), kind="blob",
         scale=7.0, pre="identity", post="relu", cam=(11, 13, 15.0, 220.0, 50.0, 4.0311), S=80, near=1.8, far=6.6, white=True),
    dict(name="sh1", dims=(12, 12, 12), deg=1, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="blob", scale=6.0,
         pre="identity", post="relu", cam=(10, 10, 13.0, 20.0, 60.0, 4.0311), S=48, near=1.8, far=6.6, white=True),
    dict(name="sh2", dims=(12, 12, 12), deg=2, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="blob", scale=6.0,
         pre="identity", post="softplus", cam=(10, 10, 13.0, 110.0, 40.0, 4.0311), S=48, near=1.8, far=6.6, white=True),
    dict(name="sh2_diffuse", dims=(12, 12, 12), deg=2, vsize=(0.25,) * 3, loc=(0, 0, 0), kind="blob", scale=6.0,
         pre="identity", post="softplus", cam=(10, 10, 13.0, 110.0, 40.0, 4.0311), S=48, near=1.8, far=6.6, white=True,
         diffuse=True),
    dict(name="sh3", dims=(10, 10, 10), deg=3, vsize=(0.3,) * 3, loc=(0, 0, 0), kind="blob", scale=6.0,
         pre="identity", post="relu", cam=(9, 9, 12.0, 260.0, 70.0, 4.0311), S=40, near=1.8, far=6.6, white=False),
    dict(name="two_samples", dims=(8, 8, 8), deg=0, vsize=(0.375,) * 3, loc=(0, 0, 0), kind="blob", scale=6.0,
         pre="identity", post="softplus", cam=(7, 9, 10.0, 0.0, 60.0, 4.0311), S=2, near=3.0, far=4.5, white=True),
    dict(name="s256_r37", dims=(24, 24, 24), deg=0, vsize=(0.125,) * 3, loc=(0, 0, 0), kind="blob", scale=12.0,
         pre="identity", post="relu", cam=(1, 37, 50.0, 35.0, 60.0, 4.0311), S=256, near=1.8, far=6.6, white=True),
    dict(name="s512_r5", dims=(24, 24, 24), deg=0, vsize=(0.125,) * 3, loc=(0, 0, 0), kind="blob", scale=12.0,
         pre="identity", post="softplus", cam=(1, 5, 8.0, 35.0, 60.0, 4.0311), S=512, near=1.8, far=6.6, white=True),
    dict(name="cube_2x2x2", dims=(2, 2, 2), deg=0, vsize=(2.0,) * 3, loc=(0, 0, 0), kind="cube", scale=1.0,
         pre="identity", post="relu", cam=(10, 10, 12.0, 90.0, 0.0, 10.0), S=128, near=5.0, far=18.0, white=True),
]


def build_case(ref, spec, seed):
    rng = np.random.default_rng(seed)
    n_feat = 3 * (spec["deg"] + 1) ** 2
    if spec["kind"] == "blob":
        dens, feat = blob_grid(spec["dims"], n_feat, rng)
    elif spec["kind"] == "uniform":
        dens, feat = uniform_grid(spec["dims"], n_feat, rng)
    else:  # the 2x2x2 +-10 cube of thre3d_reprs/tests/test_voxels.py:88-134 (values re-drawn here)
        dens = rng.uniform(-10, 10, (2, 2, 2, 1)).astype(np.float32)
        feat = (10.0 * np.sign(rng.standard_normal((2, 2, 2, 3)))).astype(np.float32)

    grid = ref["VoxelGrid"](
        densities=torch.from_numpy(dens.copy()),
        features=torch.from_numpy(feat.copy()),
        voxel_size=ref["VoxelSize"](*spec["vsize"]),
        grid_location=ref["VoxelGridLocation"](*spec["loc"]),
        density_preactivation=ACT[spec["pre"]](),
        density_postactivation=ACT[spec["post"]](),
        expected_density_scale=spec["scale"],
        tunable=True,
    )
    if spec["cam"] is not None:
        h, w, focal, yaw, pitch, radius = spec["cam"]
        rays_o, rays_d = camera_rays(ref, h, w, focal, yaw, pitch, radius)
    else:
        aabb = grid.aabb
        rays_o, rays_d = special_rays([r[0] for r in aabb], [r[1] for r in aabb])
    R, S = rays_o.shape[0], spec["S"]

    cfg = ref["SHVoxGridRenderConfig"](
        num_samples_per_ray=S,
        camera_bounds=ref["CameraBounds"](spec["near"], spec["far"]),
        perturb_sampled_points=bool(spec.get("perturb", False)),
        optimized_sampling=bool(spec.get("optimized", False)),
        linear_disparity_sampling=bool(spec.get("disparity", False)),
        white_bkgd=bool(spec["white"]),
        render_diffuse=bool(spec.get("diffuse", False)),
    )

    torch.manual_seed(seed)
    jitter = torch.rand(R, S) if cfg.perturb_sampled_points else None
    torch.manual_seed(seed)  # replay: the reference draws exactly this tensor first (sample.py:63)
    out = ref["render_sh_voxel_grid"](grid, ref["Rays"](rays_o, rays_d), cfg)

    g = torch.Generator().manual_seed(seed + 1)
    g_colour = torch.randn(R, 3, generator=g)
    acc = out.extra["accumulated_weight"]
    loss = (out.colour * g_colour).sum()
    extra = {}
    if spec.get("all_grads", False):
        g_depth = torch.randn(R, 1, generator=g)
        g_acc = torch.randn(R, 1, generator=g)
        g_disp = torch.randn(R, 1, generator=g) * (acc.detach() > 1e-3)  # disparity is NaN / ill-conditioned on misses
        disp = torch.where(acc.detach() > 1e-3, out.extra["disparity"], torch.zeros_like(acc))
        loss = loss + (out.depth * g_depth).sum() + (acc * g_acc).sum() + (disp * g_disp).sum()
        extra = dict(g_depth=g_depth.numpy(), g_acc=g_acc.numpy(), g_disp=g_disp.numpy())
    loss.backward()

    meta = dict(
        name=spec["name"], dims=list(spec["dims"]), sh_degree=spec["deg"], voxel_size=list(map(float, spec["vsize"])),
        location=list(map(float, spec["loc"])), density_scale=float(spec["scale"]), preact=spec["pre"], postact=spec["post"],
        num_samples=S, near=float(spec["near"]), far=float(spec["far"]), perturb=bool(cfg.perturb_sampled_points),
        optimized_sampling=bool(cfg.optimized_sampling), linear_disparity_sampling=bool(cfg.linear_disparity_sampling),
        white_bkgd=bool(cfg.white_bkgd), render_diffuse=bool(cfg.render_diffuse), seed=seed,
        torch_version=torch.__version__,
    )
    arrays = dict(
        densities=dens, features=feat, rays_o=rays_o.numpy(), rays_d=rays_d.numpy(),
        g_colour=g_colour.numpy(),
        colour=out.colour.detach().numpy(), depth=out.depth.detach().numpy(),
        disparity=out.extra["disparity"].detach().numpy(), accumulated_weight=acc.detach().numpy(),
        d_densities=grid.densities.grad.numpy(), d_features=grid.features.grad.numpy(),
        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
        **extra,
    )
    if jitter is not None:
        arrays["jitter"] = jitter.numpy()
    return arrays


def build_attn_case(ref, seed):
    """render_sh_voxel_grid_attn (renderers.py:108-163): 1-channel 'attn' grid, background forced to zero."""
    rng = np.random.default_rng(seed)
    dims = (14, 14, 14)
    dens, feat = blob_grid(dims, 3, rng)
    attn = (rng.standard_normal((*dims, 1)) * 2.0).astype(np.float32)
    arrays = {}
    for tag, orig in (("cur", False), ("orig", True)):
        grid = ref["VoxelGrid"](
            densities=torch.from_numpy(dens.copy()), features=torch.from_numpy(feat.copy()),
            voxel_size=ref["VoxelSize"](3 / 14, 3 / 14, 3 / 14), density_preactivation=torch.nn.Identity(),
            density_postactivation=torch.nn.ReLU(), expected_density_scale=9.0, tunable=True,
            attn=torch.from_numpy(attn.copy()),
        )
        orig_dens = (dens * 0.5 + 0.1).astype(np.float32)
        grid.orig_densities = torch.from_numpy(orig_dens.copy())
        rays_o, rays_d = camera_rays(ref, 9, 10, 12.0, 75.0, 55.0, 4.0311)
        cfg = ref["SHVoxGridRenderConfig"](
            num_samples_per_ray=56, camera_bounds=ref["CameraBounds"](1.8, 6.6), perturb_sampled_points=False, white_bkgd=True
        )
        out = ref["render_sh_voxel_grid_attn"](grid, ref["Rays"](rays_o, rays_d), cfg, None, orig)
        g = torch.Generator().manual_seed(seed + 1)
        g_attn = torch.randn(rays_o.shape[0], 1, generator=g)
        (out.attn * g_attn).sum().backward()
        arrays.update({
            f"{tag}_attn_out": out.attn.detach().numpy(), f"{tag}_depth": out.depth.detach().numpy(),
            f"{tag}_accumulated_weight": out.extra["accumulated_weight"].detach().numpy(),
            f"{tag}_d_attn": grid.attn.grad.numpy(),
            f"{tag}_d_densities": (grid.densities.grad.numpy() if grid.densities.grad is not None else np.zeros_like(dens)),
        })
    arrays.update(dict(densities=dens, features=feat, attn=attn, orig_densities=orig_dens, rays_o=rays_o.numpy(),
                       rays_d=rays_d.numpy(), g_attn=g_attn.numpy()))
    meta = dict(name="attn", dims=list(dims), voxel_size=[3 / 14] * 3, location=[0, 0, 0], density_scale=9.0,
                preact="identity", postact="relu", num_samples=56, near=1.8, far=6.6, white_bkgd=True, seed=seed)
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return arrays


def build_cameras(ref):
    arrays = {}
    combos = [(0.0, 60.0, 4.0311), (45.0, -45.0, 4.0311), (123.4, 15.0, 5.5), (359.0, 89.0, 2.0)]
    arrays["pose_args"] = np.asarray(combos, dtype=np.float64)
    for i, (yaw, pitch, radius) in enumerate(combos):
        pose = ref["pose_spherical"](yaw, pitch, radius)
        arrays[f"rot_{i}"] = pose.rotation.numpy()
        arrays[f"trans_{i}"] = pose.translation.numpy()
        rays = ref["cast_rays"](ref["CameraIntrinsics"](5, 7, 6.5), pose)
        arrays[f"rays_o_{i}"] = rays.origins.numpy()
        arrays[f"rays_d_{i}"] = rays.directions.numpy()
    poses = ref["get_thre360_animation_poses"](4.0311, 60.0, 9)
    arrays["thre360_rot"] = np.stack([p.rotation.numpy() for p in poses])
    arrays["thre360_trans"] = np.stack([p.translation.numpy() for p in poses])
    return arrays


def build_volmodel(ref, seed):
    """VolumetricModel.render: no-grad chunk loop with a partial last chunk, override kwargs, collate + reshape."""
    rng = np.random.default_rng(seed)
    dims = (12, 12, 12)
    dens, feat = blob_grid(dims, 3, rng)
    grid = ref["VoxelGrid"](
        densities=torch.from_numpy(dens.copy()), features=torch.from_numpy(feat.copy()),
        voxel_size=ref["VoxelSize"](0.25, 0.25, 0.25), density_preactivation=torch.nn.Identity(),
        density_postactivation=torch.nn.ReLU(), expected_density_scale=8.0, tunable=True,
    )
    vm = ref["VolumetricModel"](
        thre3d_repr=grid, render_procedure=ref["render_sh_voxel_grid"],
        render_config=ref["SHVoxGridRenderConfig"](num_samples_per_ray=32, camera_bounds=ref["CameraBounds"](1.8, 6.6),
                                                   white_bkgd=True, perturb_sampled_points=False),
        device=torch.device("cpu"),
    )
    pose = ref["pose_spherical"](40.0, 60.0, 4.0311)
    intr = ref["CameraIntrinsics"](9, 11, 12.0)
    out = vm.render(pose, intr, parallel_rays_chunk_size=40, num_samples_per_ray=48, optimized_sampling=True)
    meta = dict(dims=list(dims), voxel_size=[0.25] * 3, density_scale=8.0, preact="identity", postact="relu",
                height=9, width=11, focal=12.0, yaw=40.0, pitch=60.0, radius=4.0311, chunk=40, S_cfg=32, S_override=48)
    return dict(densities=dens, features=feat, colour=out.colour.numpy(), depth=out.depth.numpy(),
                disparity=out.extra["disparity"].numpy(), accumulated_weight=out.extra["accumulated_weight"].numpy(),
                meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))


def build_checkpoint(ref, seed):
    """A model file written by the reference itself (torch.save of VolumetricModel.get_save_info, trainers.py:460-469):
    pickles the reference's import paths for render_sh_voxel_grid, SHVoxGridRenderConfig, VoxelSize, CameraBounds, ..."""
    rng = np.random.default_rng(seed)
    dims = (6, 7, 8)
    dens, feat = blob_grid(dims, 3, rng)
    grid = ref["VoxelGrid"](
        densities=torch.from_numpy(dens.copy()), features=torch.from_numpy(feat.copy()),
        voxel_size=ref["VoxelSize"](0.5, 0.4, 0.3), grid_location=ref["VoxelGridLocation"](0.1, 0.0, -0.1),
        density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.Softplus(), expected_density_scale=33.333,
        tunable=True,
    )
    vm = ref["VolumetricModel"](
        thre3d_repr=grid, render_procedure=ref["render_sh_voxel_grid"],
        render_config=ref["SHVoxGridRenderConfig"](num_samples_per_ray=64, camera_bounds=ref["CameraBounds"](1.8, 6.6), white_bkgd=True),
        device=torch.device("cpu"),
    )
    info = vm.get_save_info(extra_info={"camera_bounds": ref["CameraBounds"](1.8, 6.6),
                                        "camera_intrinsics": ref["CameraIntrinsics"](100, 100, 140.0), "hemispherical_radius": 4.0311})
    torch.save(info, OUT_DIR / "reference_checkpoint.pth")
    np.savez_compressed(OUT_DIR / "reference_checkpoint_values.npz", densities=dens, features=feat)


def main():
    ref = _import_reference()
    torch.set_num_threads(1)  # deterministic reduction order for the goldens
    for i, spec in enumerate(CASES):
        arrays = build_case(ref, spec, seed=42 + i)
        np.savez_compressed(OUT_DIR / f"render_{spec['name']}.npz", **arrays)
        print(f"render_{spec['name']}: R={arrays['rays_o'].shape[0]} colour[{arrays['colour'].min():.3f},{arrays['colour'].max():.3f}] "
              f"acc mean {np.nanmean(arrays['accumulated_weight']):.3f} |d_feat|max {np.abs(arrays['d_features']).max():.3e}")
    np.savez_compressed(OUT_DIR / "attn.npz", **build_attn_case(ref, 142))
    np.savez_compressed(OUT_DIR / "cameras.npz", **build_cameras(ref))
    np.savez_compressed(OUT_DIR / "volmodel.npz", **build_volmodel(ref, 242))
    build_checkpoint(ref, 342)
    print("done")


if __name__ == "__main__":
    main()
