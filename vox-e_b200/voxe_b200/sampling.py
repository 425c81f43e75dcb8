"""Training-side ray-batch sampling on the GPU: ``voxe_sample_rays`` of the C ABI.

The reference draws every batch as ``torch.randperm(B*H*W)[:sample_size]`` over rays it has cast for all pixels of the
loaded views (thre3d_atom/rendering/volumetric/utils/misc.py:126-138, thre3d_atom/modules/trainers.py:290-313).  Here

* :func:`sample_rays_from_cameras` never materialises the per-pixel rays: one launch of ``sample_size`` threads draws
  distinct pixel indices (a keyed pseudo-random permutation evaluated on demand), generates those pixels' rays from the
  poses exactly as ``cast_rays`` does, and gathers their colours;
* :func:`sample_random_rays_and_pixels` keeps the reference's signature (ray tensors in, ray batch out) and replaces only
  the 5 M-element shuffle + three gathers by that one launch -- ``thre3d_atom...misc.sample_random_rays_and_pixels_synchronously``
  routes CUDA tensors here.

Draws are keyed by the torch CUDA generator (seed, offset; the offset is advanced), so ``torch.manual_seed`` reproduces a
run; they are another realisation of "uniform sample without replacement", not ``torch.randperm``'s numbers.  Pass
``indices=`` to replay a given selection.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from voxe_b200 import _native as nat
from voxe_b200.render_function import _require_cuda, _stream_ptr


def _rng_state(dev: torch.device, generator: Optional[torch.Generator]) -> Tuple[int, int]:
    torch.cuda.init()  # the default generators exist only once CUDA is initialised
    gen = generator if generator is not None else torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
    seed, offset = int(gen.initial_seed()) & (2**64 - 1), int(gen.get_offset())
    gen.set_offset(offset + 4)
    return seed, offset


def _launch(desc: nat.VoxeSamplerDesc, dev: torch.device, poses, src_o, src_d, pixels, indices, sample_size: int, want_rays: bool):
    lib = nat.load_library()
    out_idx = torch.empty(sample_size, dtype=torch.int64, device=dev)
    rays_o = torch.empty((sample_size, 3), dtype=torch.float32, device=dev) if want_rays else None
    rays_d = torch.empty((sample_size, 3), dtype=torch.float32, device=dev) if want_rays else None
    pix = torch.empty((sample_size, pixels.shape[1]), dtype=torch.float32, device=dev) if pixels is not None else None
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        nat.check(lib.voxe_sample_rays(desc, ptr(poses), ptr(src_o), ptr(src_d), ptr(pixels), ptr(indices), sample_size, out_idx.data_ptr(),
                                       ptr(rays_o), ptr(rays_d), ptr(pix), _stream_ptr(dev)), "voxe_sample_rays")
    return rays_o, rays_d, pix, out_idx


def _check_indices(indices: Optional[Tensor], dev: torch.device, n: int) -> Tuple[Optional[Tensor], Optional[int]]:
    if indices is None:
        return None, None
    indices = indices.to(device=dev, dtype=torch.int64).contiguous()
    if indices.dim() != 1:
        raise ValueError("indices must be a 1-D tensor of flat pixel indices")
    if indices.numel() and (int(indices.min()) < 0 or int(indices.max()) >= n):
        raise IndexError(f"pixel indices must lie in [0, {n})")
    return indices, indices.numel()


def draw_indices(num_pixels: int, sample_size: int, device, generator: Optional[torch.Generator] = None) -> Tensor:
    """``sample_size`` distinct indices in [0, num_pixels): what ``torch.randperm(num_pixels)[:sample_size]`` is to the
    reference, in O(sample_size)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("voxe_b200.sampling runs on CUDA only")
    if not 0 <= sample_size <= num_pixels:
        raise ValueError(f"cannot draw {sample_size} distinct indices out of {num_pixels}")
    seed, offset = _rng_state(dev, generator)
    desc = nat.VoxeSamplerDesc(num_pixels=int(num_pixels), rng_seed=seed, rng_offset=offset)
    return _launch(desc, dev, None, None, None, None, None, int(sample_size), False)[3]


def sample_rays_from_cameras(camera_intrinsics, poses: Tensor, pixels: Optional[Tensor], sample_size: int,
                             generator: Optional[torch.Generator] = None, indices: Optional[Tensor] = None):
    """Ray batch straight from cameras.  ``poses`` [B,3,4] = [R | t] (the dataset's matrices; trainers.py:293-295 splits them
    into ``CameraPose(pose[:, :3], pose[:, 3:])``), ``pixels`` [B*H*W, C] in ``images.permute(0,2,3,1).reshape(-1, C)``
    order (trainers.py:303-308) or None.  Returns ``(origins [k,3], directions [k,3], pixels [k,C] | None, indices [k])``
    with index = (b*H + row)*W + col -- the rows ``collate_rays([flatten_rays(cast_rays(...)) ...])`` would hold."""
    height, width, focal = camera_intrinsics
    poses = poses.detach()
    _require_cuda(poses)
    dev = poses.device
    if poses.dim() != 3 or tuple(poses.shape[1:]) != (3, 4):
        raise ValueError(f"poses must be [B, 3, 4] (got {tuple(poses.shape)})")
    n = poses.shape[0] * int(height) * int(width)
    if pixels is not None:
        _require_cuda(pixels)
        pixels = pixels.detach().contiguous()
        if pixels.dim() != 2 or pixels.shape[0] != n or pixels.device != dev:
            raise ValueError(f"pixels must be [B*H*W = {n}, C] on {dev} (got {tuple(pixels.shape)} on {pixels.device})")
    indices, k = _check_indices(indices, dev, n)
    sample_size = int(sample_size) if k is None else k
    if indices is None and not 0 <= sample_size <= n:
        raise ValueError(f"cannot draw {sample_size} distinct pixels out of {n}")
    seed, offset = (0, 0) if indices is not None else _rng_state(dev, generator)
    desc = nat.VoxeSamplerDesc(num_pixels=n, height=int(height), width=int(width), focal=float(focal),
                               pixel_channels=0 if pixels is None else int(pixels.shape[1]), rng_seed=seed, rng_offset=offset)
    return _launch(desc, dev, poses.contiguous(), None, None, pixels, indices, sample_size, True)


def sample_random_rays_and_pixels(rays_o: Tensor, rays_d: Tensor, pixels: Tensor, sample_size: int,
                                  generator: Optional[torch.Generator] = None, indices: Optional[Tensor] = None):
    """The reference signature (misc.py:126-138) on flat ray tensors [N,3] and pixels [N,C]: one launch instead of a
    ``randperm`` over N plus three gathers.  Returns ``(origins, directions, pixels, indices)``."""
    _require_cuda(rays_o, rays_d, pixels)
    dev = pixels.device
    n = pixels.shape[0]
    if rays_o.shape != (n, 3) or rays_d.shape != (n, 3) or pixels.dim() != 2:
        raise ValueError("expected flat rays [N, 3] and pixels [N, C]")
    indices, k = _check_indices(indices, dev, n)
    sample_size = int(sample_size) if k is None else k
    sample_size = min(sample_size, n)  # permutation[:sample_size] never yields more than N rows
    seed, offset = (0, 0) if indices is not None else _rng_state(dev, generator)
    desc = nat.VoxeSamplerDesc(num_pixels=n, pixel_channels=int(pixels.shape[1]), rng_seed=seed, rng_offset=offset)
    return _launch(desc, dev, None, rays_o.detach().contiguous(), rays_d.detach().contiguous(), pixels.detach().contiguous(),
                   indices, sample_size, True)
