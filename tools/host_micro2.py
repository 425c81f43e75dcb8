"""Host-side cost of the layers of one render call (run on a GPU box)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from bench import WL, frame_rays, make_grid_tensors, make_poses  # noqa: E402

bench.select_workload("cfg2")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
from thre3d_atom.modules.volumetric_model import VolumetricModel  # noqa: E402
from thre3d_atom.rendering.volumetric.render_interface import Rays  # noqa: E402
from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, _render_spec, render_sh_voxel_grid  # noqa: E402
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402
from thre3d_atom.utils.imaging_utils import CameraBounds  # noqa: E402
from voxe_b200 import render_function as rf  # noqa: E402

dens, feat = make_grid_tensors(dev)
grid = VoxelGrid(dens, feat, VoxelSize(*(w / d for w, d in zip(WL["world"], WL["dims"]))), density_preactivation=torch.nn.Identity(),
                 density_postactivation=torch.nn.ReLU(), expected_density_scale=WL["density_scale"], tunable=True)
cfg = SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]), white_bkgd=True,
                            perturb_sampled_points=True)
vm = VolumetricModel(grid, render_sh_voxel_grid, cfg, device=dev)
o, d = frame_rays(make_poses()[0], dev)
B = 4096
o0, d0 = o[20 * B:21 * B], d[20 * B:21 * B]
g0 = torch.randn(B, 3, device=dev)
rays0 = Rays(o0, d0)
ext = rf.bridge()
spec = _render_spec(cfg, 3, False, True)
gspec = grid.fused_spec()
packed = grid.packed_cache().get(gspec, grid.densities, grid.features)
scratch = grid.render_gradient_scratch().get(packed)


def t(name, fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:64s} host {1e6 * (t1 - t0) / n:7.1f} us   drained {1e6 * (t2 - t0) / n:7.1f} us", flush=True)


def bridge_call(grad):
    with torch.set_grad_enabled(grad):
        return ext.render(grid.densities, grid.features, packed, o0, d0, None, None, scratch, None, gspec.native_bytes(),
                          spec.native_bytes(), ext.MODE_DIRECT, False, None)


def fwd_bwd(fn):
    out = fn()
    c = out.colour if hasattr(out, "colour") else out[0]
    c.backward(g0)


t("slicing: o[s:s+B], d[s:s+B], g[s:s+B]", lambda: (o[B:2 * B], d[B:2 * B], g0[0:B]))
t("Rays(o, d)", lambda: Rays(o0, d0))
t("_render_spec + fused_spec + cache.get", lambda: (_render_spec(cfg, 3, False, True), grid.fused_spec(), grid.packed_cache().get(gspec, grid.densities, grid.features)))
t("bridge.render (no grad)", lambda: bridge_call(False))
t("bridge.render (grad, node built, never run)", lambda: bridge_call(True))
t("fused_render (grad)", lambda: rf.fused_render(gspec, spec, grid.densities, grid.features, o0, d0, cache=grid.packed_cache(), grad_scratch=grid.render_gradient_scratch()))
t("render_sh_voxel_grid (grad)", lambda: render_sh_voxel_grid(grid, rays0, cfg))
t("vm.render_rays (grad)", lambda: vm.render_rays(rays0))
t("bridge.render + colour.backward(g)", lambda: fwd_bwd(lambda: bridge_call(True)))
t("vm.render_rays + colour.backward(g)", lambda: fwd_bwd(lambda: vm.render_rays(rays0)))
x = torch.randn(B, 3, device=dev, requires_grad=True)
t("reference point: (x*2).backward(g) on a [4096,3] tensor", lambda: (x * 2).backward(g0))
with torch.autograd.set_multithreading_enabled(False):
    t("[single-threaded engine] bridge.render + colour.backward(g)", lambda: fwd_bwd(lambda: bridge_call(True)))
    t("[single-threaded engine] vm.render_rays + colour.backward(g)", lambda: fwd_bwd(lambda: vm.render_rays(rays0)))
    t("[single-threaded engine] (x*2).backward(g)", lambda: (x * 2).backward(g0))
