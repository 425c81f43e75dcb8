"""Grid rescale between the stages of progressive training (``scale_voxel_grid_with_required_output_size``,
voxels.py:409-447 upstream): the oracle against goldens from the executed reference (CPU), the CUDA kernel behind
``voxe_resample_grid`` against both (GPU), through the reference-named function."""
import json

import numpy as np
import pytest
import torch

from _golden import GOLDEN_DIR
from oracle.resample_oracle import resample_grid_oracle, rescaled_voxel_size


def _cases():
    z = np.load(GOLDEN_DIR / "resample.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def test_oracle_matches_the_executed_reference():
    z, meta = _cases()
    assert len(meta) >= 5
    for name, m in meta.items():
        for key in ("densities", "features"):
            for dtype in (np.float64, np.float32):
                got = resample_grid_oracle(z[f"{name}/{key}"], tuple(m["out"]), dtype)
                assert got.shape == z[f"{name}/out_{key}"].shape
                assert np.abs(got - z[f"{name}/out_{key}"]).max() <= 2e-6, (name, key, dtype)
        assert np.allclose(rescaled_voxel_size(m["voxel_size"], m["dims"], m["out"]), m["new_voxel_size"], rtol=1e-12)


def test_host_side_grids_keep_the_reference_call():
    """A grid that still lives on the host goes through torch's interpolate, as upstream (no kernel involved)."""
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize, scale_voxel_grid_with_required_output_size

    z, meta = _cases()
    m = meta["up_aniso"]
    grid = VoxelGrid(torch.from_numpy(z["up_aniso/densities"]), torch.from_numpy(z["up_aniso/features"]), VoxelSize(*m["voxel_size"]), tunable=False)
    new = scale_voxel_grid_with_required_output_size(grid, tuple(m["out"]))
    assert np.abs(new.features.numpy() - z["up_aniso/out_features"]).max() <= 1e-6
    assert np.allclose(list(new.voxel_size), m["new_voxel_size"])


@pytest.mark.gpu
def test_kernel_matches_golden_and_oracle():
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize, scale_voxel_grid_with_required_output_size
    from voxe_b200 import _native as nat

    z, meta = _cases()
    for name, m in meta.items():
        grid = VoxelGrid(torch.from_numpy(z[f"{name}/densities"]).cuda(), torch.from_numpy(z[f"{name}/features"]).cuda(),
                         VoxelSize(*m["voxel_size"]), tunable=True)
        before = nat.launch_count()
        new = scale_voxel_grid_with_required_output_size(grid, tuple(m["out"]))
        assert nat.launch_count() - before == 2  # one launch per tensor, no torch interpolate
        assert isinstance(new.features, torch.nn.Parameter) and new.features.is_contiguous() and new.densities.shape[-1] == 1
        assert np.allclose(list(new.voxel_size), m["new_voxel_size"])
        for key, got in (("densities", new.densities), ("features", new.features)):
            assert np.abs(got.detach().cpu().numpy() - z[f"{name}/out_{key}"]).max() <= 2e-6, (name, key)
    # a training-sized case against the fp64 oracle, and against torch's own CUDA interpolate
    g = torch.Generator().manual_seed(0)
    feat = torch.randn((40, 37, 45, 27), generator=g)
    grid = VoxelGrid(torch.randn((40, 37, 45, 1), generator=g).cuda(), feat.cuda(), VoxelSize(0.1, 0.1, 0.1), tunable=True)
    new = scale_voxel_grid_with_required_output_size(grid, (80, 74, 91))
    got = new.features.detach().cpu().numpy()
    assert np.abs(got - resample_grid_oracle(feat.numpy(), (80, 74, 91))).max() <= 2e-5             # fp64 truth (fp32 source coordinates differ by ~1e-6 voxel)
    assert np.abs(got - resample_grid_oracle(feat.numpy(), (80, 74, 91), np.float32)).max() <= 2e-6  # ATen's own precision
    ref = torch.nn.functional.interpolate(feat.cuda().permute(3, 0, 1, 2)[None], size=(80, 74, 91), mode="trilinear", align_corners=False)[0].permute(1, 2, 3, 0)
    assert float((new.features.detach() - ref).abs().max()) <= 2e-6
    with pytest.raises(NotImplementedError):
        scale_voxel_grid_with_required_output_size(grid, (8, 8, 8), mode="nearest")
