"""Row f3: training-side ray-batch sampling (``voxe_sample_rays``) against what the reference's per-iteration
``torch.randperm(B*H*W)[:k]`` + ``cast_rays`` + gathers mean (thre3d_atom/rendering/volumetric/utils/misc.py:12-50, 126-138;
thre3d_atom/modules/trainers.py:290-313).

Integer work is bit-exact: the drawn indices equal the CPU restatement of the keyed permutation
(``oracle/sampler_oracle.py``), are distinct and in range, and gathered rows equal the source rows.  Generated rays are
compared with ``cast_rays`` (pinned against the reference in tests/test_abi_and_api.py): origins exact, directions to 4 ulp
(the 3x3 rotation is three fused multiply-adds here, a batched matmul there).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle.sampler_oracle import Permutation


@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 17, 64, 1000, 4097])
def test_oracle_permutation_is_a_permutation(n):
    for seed, offset in ((42, 0), (2**40 + 7, 12)):
        assert sorted(Permutation(n, seed, offset).head(n).tolist()) == list(range(n))
    assert Permutation(4097, 42, 0).head(64).tolist() != Permutation(4097, 42, 4).head(64).tolist()


def test_argument_validation_without_a_gpu():
    from voxe_b200 import _native as nat

    lib = nat.load_library()
    d = nat.VoxeSamplerDesc(num_pixels=0)
    assert lib.voxe_sample_rays(d, None, None, None, None, None, 4, None, None, None, None, None) == 1
    d.num_pixels = 10
    assert lib.voxe_sample_rays(d, None, None, None, None, None, 11, None, None, None, None, None) == 1
    assert b"distinct" in lib.voxe_last_error()
    assert lib.voxe_sample_rays(d, None, None, None, None, None, 4, None, 1, None, None, None) == 1  # rays_o without rays_d
    assert lib.voxe_sample_rays(d, None, None, None, None, None, 4, None, 1, 1, None, None) == 1
    assert b"camera mode" in lib.voxe_last_error() or b"gather mode" in lib.voxe_last_error()
    d.height, d.width, d.focal = 3, 3, 1.0  # 10 is not a multiple of 9
    assert lib.voxe_sample_rays(d, 1, None, None, None, None, 4, None, 1, 1, None, None) == 1
    assert lib.voxe_sample_rays(d, None, None, None, None, None, 0, None, None, None, None, None) == 0
    assert ctypes.sizeof(nat.VoxeSamplerDesc) == 40


def test_cpu_tensors_keep_the_host_path():
    """The reference-named helper is host code for CPU tensors (as upstream); only CUDA tensors take the kernel."""
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.rendering.volumetric.utils.misc import sample_random_rays_and_pixels_synchronously
    from voxe_b200 import sampling

    n = 50
    tag = torch.arange(n, dtype=torch.float32)[:, None]
    rays, pix = sample_random_rays_and_pixels_synchronously(Rays(tag.repeat(1, 3), -tag.repeat(1, 3)), tag.repeat(1, 3), 20)
    assert rays.origins.shape == (20, 3) and len(set(pix[:, 0].tolist())) == 20
    assert torch.equal(rays.origins, pix) and torch.equal(rays.directions, -pix)
    with pytest.raises(RuntimeError, match="CUDA only"):
        sampling.draw_indices(10, 4, "cpu")


def _cameras(b=3, h=20, w=30, focal=27.5, seed=0):
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics, pose_spherical

    rng = np.random.default_rng(seed)
    poses = []
    for _ in range(b):
        p = pose_spherical(float(rng.uniform(-180, 180)), float(rng.uniform(-80, -10)), float(rng.uniform(2, 5)))
        poses.append(np.concatenate([np.asarray(p.rotation, dtype=np.float32), np.asarray(p.translation, dtype=np.float32).reshape(3, 1)], axis=1))
    return CameraIntrinsics(h, w, focal), torch.from_numpy(np.stack(poses))


def _reference_rows(intr, poses):
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, collate_rays, flatten_rays
    from thre3d_atom.utils.imaging_utils import CameraPose

    return collate_rays([flatten_rays(cast_rays(intr, CameraPose(p[:, :3], p[:, 3:]))) for p in poses])


@pytest.mark.gpu
@pytest.mark.parametrize("n,k", [(1, 1), (7, 7), (600, 600), (5000, 257), (5_120_000, 4096)])
def test_drawn_indices_match_the_restated_permutation(n, k):
    from voxe_b200 import sampling

    torch.cuda.init()
    torch.manual_seed(1234)
    gen = torch.cuda.default_generators[0]
    seed, offset = int(gen.initial_seed()), int(gen.get_offset())
    idx = sampling.draw_indices(n, k, "cuda")
    assert int(gen.get_offset()) == offset + 4
    assert idx.dtype == torch.int64 and idx.shape == (k,)
    assert torch.equal(idx.cpu(), torch.from_numpy(Permutation(n, seed, offset).head(k)))
    assert len(set(idx.tolist())) == k and 0 <= int(idx.min()) and int(idx.max()) < n
    again = sampling.draw_indices(n, k, "cuda")
    if n > 600:
        assert not torch.equal(idx, again)
    torch.manual_seed(1234)
    assert torch.equal(sampling.draw_indices(n, k, "cuda"), idx)


@pytest.mark.gpu
def test_draws_are_uniform():
    """4096 of 5.12 M (8 views of 800x800) per draw, 64 draws: every 1/64th of the index range gets its share, and the first
    drawn position is itself uniform."""
    from voxe_b200 import sampling

    n, k, draws, bins = 5_120_000, 4096, 64, 64
    torch.manual_seed(7)
    all_idx = torch.stack([sampling.draw_indices(n, k, "cuda") for _ in range(draws)]).cpu().numpy()
    counts = np.bincount((all_idx.ravel() * bins // n).astype(np.int64), minlength=bins)
    expected = draws * k / bins
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < 120.0, chi2  # 63 degrees of freedom: P(chi2 > 120) ~ 2e-5
    firsts = all_idx[:, 0] / n
    assert 0.3 < firsts.mean() < 0.7 and len(set(all_idx[:, 0].tolist())) == draws
    # consecutive positions are unrelated: lag-1 correlation of the index sequence of one draw
    x = all_idx[0].astype(np.float64)
    assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 0.06


@pytest.mark.gpu
def test_camera_mode_matches_cast_rays_rows():
    from voxe_b200 import _native as nat
    from voxe_b200 import sampling

    intr, poses = _cameras()
    want = _reference_rows(intr, poses)
    n = want.origins.shape[0]
    pixels = torch.rand(n, 3)
    before = nat.launch_count()
    o, d, pix, idx = sampling.sample_rays_from_cameras(intr, poses.cuda(), pixels.cuda(), 500)
    assert nat.launch_count() - before == 1
    idx_c = idx.cpu()
    assert len(set(idx_c.tolist())) == 500
    assert torch.equal(o.cpu(), want.origins[idx_c])
    assert torch.equal(pix.cpu(), pixels[idx_c])
    got, ref = d.cpu(), want.directions[idx_c]
    assert (got - ref).abs().max() <= 4 * 1.2e-7 * ref.abs().max()
    # every pixel exactly once when the whole set is drawn; injected indices replay a selection bit for bit
    o_all, d_all, _, idx_all = sampling.sample_rays_from_cameras(intr, poses.cuda(), None, n)
    assert sorted(idx_all.tolist()) == list(range(n))
    o2, d2, pix2, idx2 = sampling.sample_rays_from_cameras(intr, poses.cuda(), pixels.cuda(), 0, indices=idx)
    assert torch.equal(idx2, idx) and torch.equal(o2, o) and torch.equal(d2, d) and torch.equal(pix2, pix)
    with pytest.raises(IndexError):
        sampling.sample_rays_from_cameras(intr, poses.cuda(), None, 0, indices=torch.tensor([n]))


@pytest.mark.gpu
def test_reference_signature_routes_cuda_tensors_through_the_kernel():
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.rendering.volumetric.utils.misc import sample_random_rays_and_pixels_synchronously
    from voxe_b200 import _native as nat

    intr, poses = _cameras(b=2, h=33, w=17)
    rows = _reference_rows(intr, poses)
    n = rows.origins.shape[0]
    pixels = torch.cat([torch.arange(n, dtype=torch.float32)[:, None], torch.rand(n, 2)], dim=1)  # column 0 names the row
    rays_c = Rays(rows.origins.cuda(), rows.directions.cuda())
    before = nat.launch_count()
    batch, pix = sample_random_rays_and_pixels_synchronously(rays_c, pixels.cuda(), 256)
    assert nat.launch_count() - before == 1
    chosen = pix[:, 0].long().cpu()
    assert len(set(chosen.tolist())) == 256
    assert torch.equal(pix.cpu(), pixels[chosen])
    assert torch.equal(batch.origins.cpu(), rows.origins[chosen]) and torch.equal(batch.directions.cpu(), rows.directions[chosen])
    # more rows requested than exist: the whole set, once each (permutation[:k] semantics)
    batch, pix = sample_random_rays_and_pixels_synchronously(rays_c, pixels.cuda(), n + 5)
    assert sorted(pix[:, 0].long().tolist()) == list(range(n))


@pytest.mark.gpu
def test_sampled_batch_renders_like_the_gathered_rows():
    """End of the chain: a batch drawn from cameras renders to the same pixels as the same rows of cast_rays."""
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds
    from voxe_b200 import sampling

    intr, poses = _cameras(b=2, h=24, w=24, focal=30.0, seed=3)
    rows = _reference_rows(intr, poses)
    g = torch.Generator().manual_seed(0)
    grid = VoxelGrid((torch.rand((16, 16, 16, 1), generator=g) * 2 - 1).cuda(), (torch.rand((16, 16, 16, 3), generator=g) * 2 - 1).cuda(),
                     VoxelSize(3 / 16, 3 / 16, 3 / 16), density_postactivation=torch.nn.ReLU(), expected_density_scale=20.0, tunable=False)
    vm = VolumetricModel(grid, render_sh_voxel_grid, SHVoxGridRenderConfig(num_samples_per_ray=64, camera_bounds=CameraBounds(0.5, 8.0),
                                                                          perturb_sampled_points=False), device=torch.device("cuda"))
    o, d, _, idx = sampling.sample_rays_from_cameras(intr, poses.cuda(), None, 300)
    with torch.no_grad():
        ours = vm.render_rays(Rays(o, d)).colour
        ref = vm.render_rays(Rays(rows.origins[idx.cpu()].cuda(), rows.directions[idx.cpu()].cuda())).colour
    assert (ours - ref).abs().max() <= 1e-4
