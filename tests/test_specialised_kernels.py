"""The flag-specialised kernel variants (the default; VOXE_SPECIALISED_KERNELS=0 selects the generic kernels) must be
indistinguishable from the generic kernels: the render parity suite is re-run in a child process with the generic kernels
forced (the default run of the suite exercises the specialised ones), and a direct A/B on seeded cases checks pixels equal
to rounding (same arithmetic) and gradients within atomics noise."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu

AB = r"""
import sys, torch
sys.path[:0] = [r"%s", r"%s"]
from voxe_b200 import _native as nat
from thre3d_atom.modules.volumetric_model import VolumetricModel
from thre3d_atom.rendering.volumetric.render_interface import Rays
from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
from thre3d_atom.utils.imaging_utils import CameraBounds
out = {}
for deg in (0, 2):
    for post in (torch.nn.ReLU(), torch.nn.Softplus()):
        for perturb in (False, True):
            g = torch.Generator().manual_seed(deg)
            dims = (24, 20, 28)
            grid = VoxelGrid((torch.rand((*dims, 1), generator=g) * 2 - 0.7).cuda(), (torch.rand((*dims, 3 * (deg + 1) ** 2), generator=g) * 2 - 1).cuda(),
                             VoxelSize(*(3.0 / d for d in dims)), density_preactivation=torch.nn.Identity(), density_postactivation=post, expected_density_scale=12.0, tunable=True)
            vm = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(num_samples_per_ray=128, camera_bounds=CameraBounds(1.0, 7.0),
                                 white_bkgd=True, perturb_sampled_points=perturb), device=torch.device("cuda"))
            o = torch.tensor([[0.3, -3.5, 1.0]]).repeat(1000, 1).cuda()
            d = (torch.rand((1000, 3), generator=g) - 0.5).cuda() * 0.6 + torch.tensor([-0.1, 1.0, -0.3]).cuda()
            torch.manual_seed(5)
            res = vm.render_rays(Rays(o, d))
            (res.colour * torch.linspace(-1, 1, 3000).reshape(1000, 3).cuda()).sum().backward()
            out[(deg, type(post).__name__, perturb)] = (res.colour.detach().cpu(), res.depth.detach().cpu(), grid.densities.grad.cpu(), grid.features.grad.cpu())
torch.save({"out": out, "specialised": int(nat.load_library().voxe_specialised_launch_count())}, sys.argv[1])
""" % (ROOT, ROOT / "vox-e_b200")


def _child(env_value, *args):
    env = dict(os.environ, VOXE_SPECIALISED_KERNELS=env_value)
    return subprocess.run([sys.executable, *args], env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_parity_suite_with_generic_kernels():
    r = _child("0", "-m", "pytest", "tests/test_cuda_parity.py", "tests/test_grad_handover.py", "tests/test_kernel_jitter.py",
               "tests/test_baseline_configs.py", "-q", "-m", "gpu", "-x")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_specialised_equals_generic(tmp_path):
    import torch

    got = {}
    for flag in ("0", "1"):
        path = tmp_path / f"ab{flag}.pt"
        r = _child(flag, "-c", AB, str(path))
        assert r.returncode == 0, r.stderr[-3000:]
        got[flag] = torch.load(path)
    assert got["0"]["specialised"] == 0 and got["1"]["specialised"] == 2 * 8  # every forward and backward launch
    for key, (c0, z0, gd0, gf0) in got["0"]["out"].items():
        c1, z1, gd1, gf1 = got["1"]["out"][key]
        # same arithmetic; the compiler may still contract a multiply-add differently in the two variants
        assert (c0 - c1).abs().max().item() <= 2e-6 and (z0 - z1).abs().max().item() <= 2e-5, key
        for a, b in ((gd0, gd1), (gf0, gf1)):
            assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-30), key  # atomics order only
