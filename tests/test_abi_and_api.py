"""CPU-side checks: the C-ABI library loads and exports every symbol ``include/voxe.h`` declares, struct layouts agree
between the header and the ctypes binding, the reference-facing API keeps the reference's names / shapes / error
behaviour, camera helpers match the reference's golden values, a checkpoint written by the reference loads, and the
product path refuses to run without CUDA (there is no CPU fallback)."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from _golden import GOLDEN_DIR, load_npz

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "voxe.h"


def _declared_symbols():
    text = HEADER.read_text()
    return re.findall(r"VOXE_API\s+[\w\s\*]+?\b(voxe_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    from voxe_b200 import _native as nat

    lib = nat.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 10
    assert set(declared) == set(nat.EXPORTS), "binding and header disagree about the ABI surface"
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.voxe_abi_version() == nat.ABI_VERSION
    assert [lib.voxe_packed_channels(f) for f in (1, 3, 12, 27, 48)] == [4, 4, 16, 28, 52]


def test_struct_layouts_match_the_header(tmp_path):
    from voxe_b200 import _native as nat

    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "voxe.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(VoxeGridDesc), offsetof(VoxeGridDesc, aabb_lo),'
        " offsetof(VoxeGridDesc, density_scale), sizeof(VoxeRenderDesc), offsetof(VoxeRenderDesc, flags),"
        " offsetof(VoxeRenderDesc, noise_std), offsetof(VoxeRenderDesc, rng_offset), sizeof(VoxeCameraDesc),"
        " offsetof(VoxeCameraDesc, translation), sizeof(VoxeAdamDesc), sizeof(VoxeSamplerDesc), offsetof(VoxeSamplerDesc, rng_seed));return 0;}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)  # header is plain C
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [ctypes.sizeof(nat.VoxeGridDesc), nat.VoxeGridDesc.aabb_lo.offset, nat.VoxeGridDesc.density_scale.offset,
            ctypes.sizeof(nat.VoxeRenderDesc), nat.VoxeRenderDesc.flags.offset, nat.VoxeRenderDesc.noise_std.offset,
            nat.VoxeRenderDesc.rng_offset.offset, ctypes.sizeof(nat.VoxeCameraDesc), nat.VoxeCameraDesc.translation.offset,
            ctypes.sizeof(nat.VoxeAdamDesc), ctypes.sizeof(nat.VoxeSamplerDesc), nat.VoxeSamplerDesc.rng_seed.offset]
    assert got == want


def test_argument_validation_without_a_gpu():
    """Descriptor checks run before any launch, so they are testable on the CPU box."""
    from voxe_b200 import _native as nat

    lib = nat.load_library()
    gd, rd = nat.VoxeGridDesc(), nat.VoxeRenderDesc()
    assert lib.voxe_render_fwd(gd, rd, None, None, None, None, None, None, None, None, None, None, 16, None) == 1
    assert b"dims" in lib.voxe_last_error()
    for a in range(3):
        gd.dims[a] = 8
    gd.n_features, gd.channels = 3, 4
    rd.num_samples, rd.sh_degree, rd.n_colour = 1, 0, 3
    assert lib.voxe_render_fwd(gd, rd, None, None, None, None, None, None, None, None, None, None, 16, None) == 1
    assert b"num_samples" in lib.voxe_last_error()
    rd.num_samples, rd.sh_degree = 64, 4
    assert lib.voxe_render_fwd(gd, rd, None, None, None, None, None, None, None, None, None, None, 16, None) == 2  # unsupported
    rd.sh_degree, rd.flags = 0, nat.FLAG_PERTURB  # jitter == NULL is legal (in-kernel draws); the buffers are still checked
    assert lib.voxe_render_fwd(gd, rd, None, None, None, None, None, None, None, None, None, None, 16, None) == 1
    assert b"NULL buffer" in lib.voxe_last_error()
    rd.flags = 0
    # per segment (8 segments of 8 samples at S=64): (n_colour+3) summary floats + one float4 per sample slot
    assert lib.voxe_saved_floats(rd, 4096) == (6 + 4 * 8) * 8 * 4096
    assert lib.voxe_set_tuning(3, 33, 0) == 1 and lib.voxe_set_tuning(3, 5, 70) == 1 and lib.voxe_set_tuning(0, 0, 0) == 0
    # the collective's descriptor checks (no launch without a GPU)
    peers = nat.VoxePeerDesc()
    peers.world_size, peers.rank = 2, 2
    assert lib.voxe_allreduce_grads_peer(peers, 1024, None, None) == 1 and b"rank" in lib.voxe_last_error()
    peers.rank = 0
    assert lib.voxe_allreduce_grads_peer(peers, 1024, None, None) == 1 and b"NULL buffer" in lib.voxe_last_error()
    assert lib.voxe_allreduce_grads_peer(peers, 1022, None, None) == 1 and b"multiple of 4" in lib.voxe_last_error()
    assert lib.voxe_allreduce_grads_peer_sparse(peers, gd, None, 0, None, None) == 1 and b"NULL buffer" in lib.voxe_last_error()
    assert lib.voxe_peer_touched_bytes(gd) == 128 and lib.voxe_query_points(gd, None, None, None, 8, None) == 1
    assert lib.voxe_allreduce_grads(None, None, 16, None) == 1
    assert lib.voxe_touched_bytes(gd) == 5 * 5 * 5 and lib.voxe_consume_grad(gd, None, None, None, None, 0, None) == 1
    dims_in, dims_out = (ctypes.c_int32 * 3)(4, 4, 4), (ctypes.c_int32 * 3)(0, 4, 4)
    assert lib.voxe_resample_grid(None, dims_in, 3, None, dims_out, None) == 1
    with pytest.raises(NotImplementedError):
        nat.check(2, "x")
    with pytest.raises(nat.NativeLibraryError):
        nat.check(1, "x")


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from voxe_b200 import _native as nat

    monkeypatch.setenv("VOXE_LIBRARY", str(tmp_path / "nope.so"))
    monkeypatch.setattr(nat, "_lib", None)
    with pytest.raises(nat.NativeLibraryError, match="no CPU"):
        nat.load_library()


def test_render_refuses_cpu_tensors_and_unfused_callables():
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds

    grid = VoxelGrid(torch.zeros(4, 4, 4, 1), torch.zeros(4, 4, 4, 3), VoxelSize(1, 1, 1))
    rays = Rays(torch.zeros(5, 3), torch.ones(5, 3))
    cfg = SHVoxGridRenderConfig(num_samples_per_ray=8, camera_bounds=CameraBounds(1.0, 2.0))
    with pytest.raises(RuntimeError, match="CUDA only"):
        render_sh_voxel_grid(grid, rays, cfg)
    with pytest.raises(AssertionError):  # render interface only takes flat rays (render_interface.py:164-166 upstream)
        render_sh_voxel_grid(grid, Rays(torch.zeros(2, 2, 3), torch.ones(2, 2, 3)), cfg)
    cfg.radiance_hdr_tone_map = torch.tanh
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(grid, rays, cfg)
    cfg.radiance_hdr_tone_map = torch.sigmoid
    bad = VoxelGrid(torch.zeros(4, 4, 4, 1), torch.zeros(4, 4, 4, 3), VoxelSize(1, 1, 1), density_postactivation=torch.nn.Tanh())
    with pytest.raises(NotImplementedError):
        render_sh_voxel_grid(bad, rays, cfg)
    with pytest.raises(RuntimeError, match="CUDA only"):
        grid(torch.zeros(3, 3))  # the stand-alone point query is a CUDA kernel too


def test_value_types_keep_the_reference_contract():
    from thre3d_atom.rendering.volumetric.render_interface import Rays, RenderOut, RenderOutAttn

    rays = Rays(torch.arange(30.0).reshape(10, 3), torch.ones(10, 3))
    assert len(rays) == 10 and len(rays[2:5]) == 3 and len(rays[torch.tensor([1, 3])]) == 2
    assert torch.equal(rays[4:6].origins, rays.origins[4:6])
    with pytest.raises(AssertionError):
        Rays(torch.zeros(3, 3), torch.zeros(4, 3))
    with pytest.raises(AssertionError):
        Rays(torch.zeros(3, 2), torch.zeros(3, 2))
    out = RenderOut(colour=torch.zeros(4, 3, requires_grad=True), depth=torch.zeros(4, 1))
    assert out.extra == {} and not out.detach().colour.requires_grad
    with pytest.raises(AssertionError):
        RenderOut(colour=torch.zeros(4, 4), depth=torch.zeros(4, 1))
    with pytest.raises(AssertionError):
        RenderOut(colour=torch.zeros(4, 3), depth=torch.zeros(4, 2))
    attn = RenderOutAttn(attn=torch.zeros(4, 1), depth=torch.zeros(4, 1), extra={"k": torch.ones(4, 1)})
    assert attn.to(torch.device("cpu")).extra["k"].shape == (4, 1)


def test_voxel_grid_surface():
    from thre3d_atom.thre3d_reprs.voxels import (VoxelGrid, VoxelGridLocation, VoxelSize,
                                                  scale_voxel_grid_with_required_output_size)

    g = VoxelGrid(torch.rand(4, 6, 8, 1), torch.rand(4, 6, 8, 3), VoxelSize(0.5, 0.25, 0.125), VoxelGridLocation(1.0, 0.0, -1.0),
                  density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(), tunable=True)
    assert g.grid_dims == (4, 6, 8)
    assert g.aabb.x_range == (0.0, 2.0) and g.aabb.y_range == (-0.75, 0.75) and g.aabb.z_range == (-1.5, -0.5)
    assert set(g.state_dict()) == {"_densities", "_features"} and len(list(g.parameters())) == 2
    inside = g.test_inside_volume(torch.tensor([[1.0, 0.0, -1.0], [0.0, 0.0, -1.0], [2.5, 0.0, -1.0]]))
    assert inside.tolist() == [[True], [False], [False]]  # strict inequalities: a point on the face is outside
    assert g.get_bounding_volume_vertices().shape == (8, 3)
    with pytest.raises(AssertionError):
        g.densities = torch.zeros(4, 6, 8, 2)
    g.features = torch.zeros(4, 6, 8, 3)
    assert isinstance(g.features, torch.nn.Parameter)
    cfg = g.get_save_config_dict()
    assert cfg["voxel_size"] == VoxelSize(0.5, 0.25, 0.125) and cfg["tunable"] is True
    spec = g.fused_spec()
    assert (spec.preact, spec.postact, spec.channels, spec.n_features) == (0, 1, 4, 3)
    up = scale_voxel_grid_with_required_output_size(g, (8, 12, 16))
    assert up.grid_dims == (8, 12, 16) and up.aabb == g.aabb and up.voxel_size == VoxelSize(0.25, 0.125, 0.0625)
    with pytest.raises(AssertionError):
        VoxelGrid(torch.rand(4, 6, 8), torch.rand(4, 6, 8, 3), VoxelSize())


def test_cameras_match_reference_golden():
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, compute_expected_density_scale_for_relu_field_grid, flatten_rays
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics, get_thre360_animation_poses, pose_spherical

    _, a = load_npz("cameras")
    for i, (yaw, pitch, radius) in enumerate(a["pose_args"].tolist()):
        pose = pose_spherical(yaw, pitch, radius)
        assert torch.equal(pose.rotation, a[f"rot_{i}"]) and torch.equal(pose.translation, a[f"trans_{i}"])
        rays = cast_rays(CameraIntrinsics(5, 7, 6.5), pose)
        assert rays.origins.shape == (5, 7, 3)
        assert torch.equal(rays.origins, a[f"rays_o_{i}"]) and torch.equal(rays.directions, a[f"rays_d_{i}"])
        assert flatten_rays(rays).directions.shape == (35, 3)
    poses = get_thre360_animation_poses(4.0311, 60.0, 9)
    assert len(poses) == 8
    assert torch.equal(torch.stack([p.rotation for p in poses]), a["thre360_rot"])
    assert abs(compute_expected_density_scale_for_relu_field_grid((3.0, 3.0, 3.0)) - 33.3333333) < 1e-5


def test_oracle_camera_helpers_match_reference_golden():
    from oracle.voxe_oracle import cast_rays_np, pose_spherical_np

    _, a = load_npz("cameras")
    for i, (yaw, pitch, radius) in enumerate(a["pose_args"].tolist()):
        rot, trans = pose_spherical_np(yaw, pitch, radius)
        assert np.allclose(rot, a[f"rot_{i}"].numpy(), atol=1e-6) and np.allclose(trans, a[f"trans_{i}"].numpy(), atol=1e-6)
        o, d = cast_rays_np(5, 7, 6.5, a[f"rot_{i}"].numpy(), a[f"trans_{i}"].numpy())
        assert torch.allclose(d, a[f"rays_d_{i}"].reshape(-1, 3), atol=1e-6) and torch.allclose(o, a[f"rays_o_{i}"].reshape(-1, 3))


def test_checkpoint_written_by_the_reference_loads(tmp_path):
    """The pickled import paths (render procedure, config class, NamedTuples, activations) resolve to this package."""
    from thre3d_atom.modules.volumetric_model import VolumetricModel, create_volumetric_model_from_saved_model
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelSize, create_voxel_grid_from_saved_info_dict
    from thre3d_atom.utils.imaging_utils import CameraBounds

    vm, extra = create_volumetric_model_from_saved_model(GOLDEN_DIR / "reference_checkpoint.pth", create_voxel_grid_from_saved_info_dict)
    assert vm.render_procedure is render_sh_voxel_grid  # identity asserts of trainers.py:127-129 keep working
    assert isinstance(vm.render_config, SHVoxGridRenderConfig) and vm.render_config.camera_bounds == CameraBounds(1.8, 6.6)
    assert extra["hemispherical_radius"] == 4.0311
    vals = np.load(GOLDEN_DIR / "reference_checkpoint_values.npz")
    grid = vm.thre3d_repr
    assert grid.voxel_size == VoxelSize(0.5, 0.4, 0.3) and grid.grid_dims == (6, 7, 8)
    assert np.array_equal(grid.densities.detach().numpy(), vals["densities"])
    assert np.array_equal(grid.features.detach().numpy(), vals["features"])
    assert grid.fused_spec().postact == 2  # Softplus survived the pickle and is recognised
    # and back: what we save has the same layout and reloads
    path = tmp_path / "roundtrip.pth"
    torch.save(vm.get_save_info(extra_info=extra), path)
    vm2, _ = create_volumetric_model_from_saved_model(path, create_voxel_grid_from_saved_info_dict)
    assert torch.equal(vm2.thre3d_repr.features, grid.features)
    with pytest.raises(ValueError):
        VolumetricModel._update_render_config(vm.render_config, {"not_a_field": 1})
    cfg2 = VolumetricModel._update_render_config(vm.render_config, {"render_diffuse": True, "num_samples_per_ray": 7})
    assert cfg2.render_diffuse and cfg2.num_samples_per_ray == 7 and not vm.render_config.render_diffuse


def test_volumetric_model_chunk_loop_with_a_stub_procedure():
    """render(): no-grad, partial last chunk, kwargs reach the procedure, outputs come back [H, W, .]."""
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.render_interface import RenderOut
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics, pose_spherical

    calls = []

    def procedure(grid, rays, cfg, chunk):
        calls.append((len(rays), cfg.num_samples_per_ray, torch.is_grad_enabled()))
        idx = rays.directions[:, :1]
        return RenderOut(colour=idx.repeat(1, 3), depth=idx, extra={"disparity": idx, "accumulated_weight": idx})

    grid = VoxelGrid(torch.zeros(2, 2, 2, 1), torch.zeros(2, 2, 2, 3), VoxelSize(1, 1, 1))
    vm = VolumetricModel(grid, procedure, SHVoxGridRenderConfig(num_samples_per_ray=5, camera_bounds=CameraBounds(1, 2)), device=torch.device("cpu"))
    out = vm.render(pose_spherical(10, 20, 3.0), CameraIntrinsics(6, 7, 5.0), parallel_rays_chunk_size=16, num_samples_per_ray=9, gpu_render=False)
    assert [c[0] for c in calls] == [16, 16, 10] and all(c[1] == 9 and c[2] is False for c in calls)
    assert out.colour.shape == (6, 7, 3) and out.extra["disparity"].shape == (6, 7, 1)
    assert vm.render_config.num_samples_per_ray == 5


def test_batchify_and_misc():
    from thre3d_atom.utils.misc import batchify, check_power_of_2, compute_thre3d_grid_sizes

    double = batchify(lambda x: x * 2, collate_fn=lambda parts: torch.cat(parts), chunk_size=4)
    assert torch.equal(double(torch.arange(10)), torch.arange(10) * 2)
    assert batchify(abs, None, None) is abs
    assert compute_thre3d_grid_sizes((160, 160, 160), 4, 2.0) == [(20, 20, 20), (40, 40, 40), (80, 80, 80), (160, 160, 160)]
    assert check_power_of_2(64) and not check_power_of_2(48)


def test_product_tree_never_touches_the_oracle_or_a_cpu_path():
    """The oracle is test infrastructure: nothing under vox-e_b200/ may import it, and the only places outside tests/ that
    do are the two the contract names (smoke() in __graft_entry__.py, the CPU legs of bench.py).  The product's native
    sources must not contain host-side compute fallbacks either: every kernel entry point goes through a launch."""
    import ast

    offenders = []
    for path in (ROOT / "vox-e_b200").rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            offenders += [f"{path.relative_to(ROOT)} imports {n}" for n in names if n.split(".")[0] == "oracle"]
    assert not offenders, offenders
    allowed = {"bench.py", "__graft_entry__.py"}
    for path in ROOT.glob("*.py"):
        if "oracle" in path.read_text() and path.name not in allowed:
            offenders.append(path.name)
    for path in (ROOT / "tools").glob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                offenders.append(f"tools/{path.name}")
    assert not offenders, offenders
    # a missing library is an error, not a detour
    from voxe_b200 import _native as nat

    src = (ROOT / "vox-e_b200" / "voxe_b200" / "_native.py").read_text()
    assert "raise NativeLibraryError" in src and nat.NativeLibraryError.__mro__[1] is RuntimeError
