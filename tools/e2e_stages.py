"""Per-stage host time of the e2e loop of bench.py (run on a GPU box): where does a 4096-ray batch spend its 300 us?"""
import sys
import time
from collections import defaultdict
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
from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid  # noqa: E402
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402
from thre3d_atom.utils.imaging_utils import CameraBounds  # noqa: E402

deferred = "deferred" in sys.argv
sync_stage = "sync" in sys.argv
dens, feat = make_grid_tensors(dev)
grid = VoxelGrid(dens, feat, VoxelSize(*(w / d for w, d in zip(WL["world"], WL["dims"]))), density_preactivation=torch.nn.Identity(),
                 density_postactivation=torch.nn.ReLU(), expected_density_scale=WL["density_scale"], tunable=True)
vm = VolumetricModel(grid, render_sh_voxel_grid,
                     SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]),
                                           white_bkgd=True, perturb_sampled_points=True), device=dev)
if deferred:
    grid.accumulate_render_gradients()
poses = make_poses()
g = torch.Generator().manual_seed(7)
host = []
for p in poses[:3]:
    o, d = frame_rays(p, torch.device("cpu"))
    host.append((o.pin_memory(), d.pin_memory(), torch.randn(o.shape[0], 3, generator=g).pin_memory()))
R, B = host[0][0].shape[0], WL["batch"]
colour_host = torch.empty(R, 3).pin_memory()
acc = defaultdict(float)


def mark(name, t0):
    if sync_stage:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    acc[name] += t1 - t0
    return t1


def one_frame(idx):
    o_h, d_h, g_h = host[idx % len(host)]
    grid.densities.grad = None
    grid.features.grad = None
    t = time.perf_counter()
    o = o_h.to(dev, non_blocking=True)
    d = d_h.to(dev, non_blocking=True)
    gc = g_h.to(dev, non_blocking=True)
    t = mark("h2d(frame)", t)
    colours = []
    for s in range(0, R, B):
        t = time.perf_counter()
        rays = Rays(o[s:s + B], d[s:s + B])
        gslice = gc[s:s + B]
        t = mark("slice", t)
        out = vm.render_rays(rays)
        t = mark("render_rays", t)
        out.colour.backward(gslice)
        t = mark("backward", t)
        colours.append(out.colour.detach())
        t = mark("collect", t)
    t = time.perf_counter()
    colour = torch.cat(colours)
    loss_total = (colour * gc).sum()
    colour_host.copy_(colour, non_blocking=True)
    if deferred:
        grid.materialize_render_gradients()
    t = mark("frame tail", t)
    v = float(loss_total.item())
    mark("item", t)
    return v


if "nothreads" in sys.argv:  # backward on the calling thread (the stock caller-side switch)
    torch.autograd.set_multithreading_enabled(False)
for k in range(3):
    one_frame(k)
acc.clear()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
N = 6
for k in range(N):
    one_frame(k)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
nb = N * ((R + B - 1) // B)
print(f"mode={'deferred' if deferred else 'default'} sync_stage={sync_stage}: wall {1e3 * wall / N:.2f} ms/frame, device events {e0.elapsed_time(e1) / N:.2f} ms/frame, "
      f"{1e6 * wall / nb:.1f} us/batch")
for k, v in acc.items():
    print(f"  {k:14s} {1e6 * v / nb:8.1f} us/batch")
