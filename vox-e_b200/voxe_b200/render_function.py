"""Host side of one render call: descriptors, packed-volume cache, gradient volumes, and the hand-off to the C++
autograd node (``voxe_b200._voxe_torch``, csrc/voxe_torch.cpp) that calls ``voxe_render_fwd`` / ``voxe_render_bwd``.

``fused_render`` is what ``thre3d_atom.thre3d_reprs.renderers.render_sh_voxel_grid`` calls: it replaces the reference's
sampler -> point processor -> accumulator chain (render_interface.py:140-171) and its autograd graph with one
forward kernel and one backward kernel linked by a 16-byte-per-sample workspace (the reference retains ~35 floats per
sample in autograd).

Gradients w.r.t. ``densities`` [X,Y,Z,1] and ``features`` [X,Y,Z,F] reach ``.grad`` as dense tensors shaped like the
parameters, as autograd would have produced (see ``DIRECT_GRAD_ACCUMULATION`` for how).  Rays never receive gradients
(they never require them upstream: ``cast_rays`` builds them from poses, misc.py:30-50).
"""
from __future__ import annotations

import dataclasses
import functools
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from voxe_b200 import _native as nat

_bridge_module = None


def bridge():
    """The C++ torch bridge (autograd node).  Like the C-ABI library it has no fallback: a missing build is an error."""
    global _bridge_module
    if _bridge_module is None:
        nat.load_library()  # the bridge links against libvoxe_sm100a.so; surface a missing / stale library first
        try:
            from voxe_b200 import _voxe_torch
        except ImportError as exc:
            raise nat.NativeLibraryError(
                f"cannot import voxe_b200._voxe_torch ({exc}): build it with `make -C vox-e_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`).  There is no Python fallback for the render path."
            ) from exc
        if _voxe_torch.abi_version() != nat.ABI_VERSION:
            raise nat.NativeLibraryError("voxe_b200._voxe_torch was built against another ABI version; rebuild")
        _bridge_module = _voxe_torch
    return _bridge_module


@dataclasses.dataclass(frozen=True)
class FusedGridSpec:
    """Geometry + activations of a voxel grid in the form the kernels take (VoxeGridDesc of include/voxe.h)."""

    dims: Tuple[int, int, int]
    n_features: int
    aabb: Tuple[Tuple[float, float], Tuple[float, float], Tuple[float, float]]  # python doubles, voxels.py:198-223
    density_scale: float
    preact: int
    postact: int

    @property
    def channels(self) -> int:
        return ((self.n_features + 1 + 3) // 4) * 4

    def __post_init__(self) -> None:  # the ctypes struct and its bytes are built once per description object
        native = self._build_native()
        object.__setattr__(self, "_native", native)
        object.__setattr__(self, "_native_bytes", bytes(native))

    def to_native(self) -> nat.VoxeGridDesc:
        return self._native

    def native_bytes(self) -> bytes:
        return self._native_bytes

    def _build_native(self) -> nat.VoxeGridDesc:
        d = nat.VoxeGridDesc()
        for a in range(3):
            lo32, hi32 = np.float32(self.aabb[a][0]), np.float32(self.aabb[a][1])
            # adjust_dynamic_range(..., drange_out=(-1, 1), slack=True): numpy fp32 arithmetic (imaging_utils.py:57-63)
            scale = (np.float32(1.0) - np.float32(-1.0)) / (hi32 - lo32)
            bias = np.float32(-1.0) - lo32 * scale
            d.dims[a] = int(self.dims[a])
            d.aabb_lo[a], d.aabb_hi[a] = float(lo32), float(hi32)
            d.norm_scale[a], d.norm_bias[a] = float(scale), float(bias)
        d.n_features = int(self.n_features)
        d.channels = int(self.channels)
        d.density_scale = float(self.density_scale)
        d.preact, d.postact = int(self.preact), int(self.postact)
        return d


@dataclasses.dataclass(frozen=True)
class FusedRenderSpec:
    """One render call (VoxeRenderDesc of include/voxe.h)."""

    num_samples: int
    near: float
    far: float
    flags: int
    sh_degree: int
    n_colour: int
    noise_std: float = 0.0

    def __post_init__(self) -> None:
        native = self._build_native()
        object.__setattr__(self, "_native", native)
        object.__setattr__(self, "_native_bytes", bytes(native))

    def to_native(self) -> nat.VoxeRenderDesc:
        return self._native

    def native_bytes(self) -> bytes:
        return self._native_bytes

    def _build_native(self) -> nat.VoxeRenderDesc:
        r = nat.VoxeRenderDesc()
        r.num_samples, r.near, r.far = int(self.num_samples), float(self.near), float(self.far)
        r.flags, r.sh_degree, r.n_colour, r.noise_std = int(self.flags), int(self.sh_degree), int(self.n_colour), float(self.noise_std)
        return r


# The reference draws ``torch.randn(R, S)`` for the density noise on EVERY call, even when
# stochastic_density_noise_std == 0 and the draw is multiplied away (accumulate.py:59-62).  Skipping that draw changes
# nothing in one call's result but leaves the global generator in a different state for the next call.  Set this to True
# to consume the generator exactly like the reference (one extra RNG kernel per call) when replaying its seeded runs.
STRICT_REFERENCE_RNG = False

# How the voxel gradients of a plain ``loss.backward()`` reach ``.grad``.  The backward kernel scatters into the grid's
# packed gradient volume.  True (default): the autograd node then adds the touched voxels straight into
# ``_densities.grad`` / ``_features.grad`` (creating zero gradients on first use) with one sparse pass
# (``voxe_consume_grad``) -- AccumulateGrad's job without its three full-grid passes per call.  The node falls back to
# returning dense gradients to autograd by itself whenever that would not be equivalent: ``torch.autograd.grad`` /
# ``backward(inputs=...)``, parameters with tensor or post-accumulate hooks, non-leaf or oddly laid out ``.grad``.
# False: always return dense gradients to autograd.
DIRECT_GRAD_ACCUMULATION = True


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*tensors: Tensor) -> torch.device:
    dev = tensors[0].device
    if dev.type != "cuda":
        raise RuntimeError(
            "the fused Vox-E render path runs on CUDA only (tensors are on "
            f"'{dev}'); there is deliberately no CPU fallback in this package"
        )
    for t in tensors:
        if t.device != dev:
            raise RuntimeError(f"all render inputs must live on one device (got {t.device} and {dev})")
        if t.dtype != torch.float32:
            raise TypeError(f"the render path computes in fp32 (got {t.dtype})")
    return dev


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def pack_volume(spec: FusedGridSpec, densities: Tensor, features: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """Packed volume (2x2x2 bricks of concat(features, densities, 0-padding) voxels) via ``voxe_pack_grid``: a flat
    fp32 tensor whose layout is private to the library."""
    dev = _require_cuda(densities, features)
    lib = nat.load_library()
    dens, feat = densities.detach().contiguous(), features.detach().contiguous()
    gd = spec.to_native()
    n = int(lib.voxe_packed_floats(gd))
    if out is None or out.numel() != n or out.device != dev:
        out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.voxe_pack_grid(gd, dens.data_ptr(), feat.data_ptr(), out.data_ptr(), _stream_ptr(dev)), "voxe_pack_grid")
    return out


class PackedVolumeCache:
    """Keeps the packed volume of a (densities, features) pair until either tensor is replaced or written to.

    Callers of the reference API both mutate parameters in place (optimiser steps, ``.data`` writes) and replace
    them wholesale (voxels.py:145-163 setters; attn_grid_trainer.py:546-550), so staleness is detected from
    ``data_ptr`` + ``_version`` of both tensors on every call.  ``.data`` writes do not bump ``_version``: callers that
    do that must call :meth:`invalidate`.
    """

    def __init__(self) -> None:
        self._key = None
        self._packed: Optional[Tensor] = None

    def invalidate(self) -> None:
        self._key = None

    def mark_fresh(self, spec: FusedGridSpec, densities: Tensor, features: Tensor) -> None:
        """The packed volume was brought up to date out of band (the fused optimiser step rewrites it)."""
        if self._packed is not None:
            self._key = (densities.data_ptr(), densities._version, features.data_ptr(), features._version, densities.device, spec)

    def peek(self) -> Optional[Tensor]:
        return self._packed

    def get(self, spec: FusedGridSpec, densities: Tensor, features: Tensor) -> Tensor:
        key = (densities.data_ptr(), densities._version, features.data_ptr(), features._version, densities.device, spec)
        if self._packed is None or key != self._key:
            # Repack in place only while nobody else holds the buffer.  An autograd graph that is still alive saved it for
            # its backward (a render of other tensors through the same cache -- render_rays_attn with and without
            # orig_densities -- or of a parameter that has been replaced since): that graph keeps the old buffer and this
            # call gets a fresh one.  (Python wrapper = 1 use; every live graph node adds one.)
            reuse = self._packed if self._packed is not None and self._packed._use_count() <= 1 else None
            self._packed = pack_volume(spec, densities, features, out=reuse)
            self._key = key
        return self._packed


class PackedGradAccumulator:
    """Persistent packed gradient volume of one grid ("deferred gradients").

    With an accumulator attached, the backward kernel scatter-adds straight into this buffer and autograd receives no
    gradient for ``_densities`` / ``_features``: the per-call zero-fill, unpack and ``.grad +=`` passes (3 x 65 MB at
    160^3) disappear from every render call.  The gradients become visible to torch when :meth:`materialize` splits the
    buffer into ``.grad`` (done automatically right before any ``torch.optim`` step, see ``voxe_b200.optim``), or are
    consumed in place by ``FusedVoxelAdam``.  Opt-in, because code that reads ``.grad`` between ``backward()`` and the
    optimiser step would not see the render's contribution."""

    def __init__(self) -> None:
        self.buffer: Optional[Tensor] = None
        self.flag = torch.zeros(1, dtype=torch.int64)  # set by the C++ backward node when it scatters into the buffer
        # trail of the backward's scatter (one byte per 2x2x2 brick, include/voxe.h: voxe_render_bwd) and the last tag
        # handed out for it; used by the sparse direct hand-over, ignored by the deferred mode
        self.touched: Optional[Tensor] = None
        self.touch_tag = torch.zeros(1, dtype=torch.int64)
        # deferred mode on a large grid: keep the trail as well, so that the step's exchange (PeerGradVolume.allreduce_sparse)
        # and the hand-over into .grad follow the bricks the step touched instead of sweeping the volume (enable_trail)
        self.sparse_sink = False
        self.peer_volume = None  # set by PeerGradVolume.adopt: the volume lives in peer-mapped memory

    @property
    def dirty(self) -> bool:
        return bool(self.flag[0] != 0)

    @dirty.setter
    def dirty(self, value: bool) -> None:
        self.flag[0] = 1 if value else 0

    def get(self, like: Tensor) -> Tensor:
        if self.buffer is None or self.buffer.numel() != like.numel() or self.buffer.device != like.device:
            self.buffer = torch.zeros_like(like)
            self.touched = None
            self.dirty = False
        return self.buffer

    def get_touched(self, spec: "FusedGridSpec") -> Tensor:
        if self.touched is None or self.touched.device != self.buffer.device:
            n = int(nat.load_library().voxe_touched_bytes(spec.to_native()))
            self.touched = torch.zeros(n, dtype=torch.uint8, device=self.buffer.device)
        return self.touched

    def enable_trail(self, spec: "FusedGridSpec", like: Tensor, touched: Optional[Tensor] = None) -> None:
        """Deferred gradients with a brick-flag trail: every backward of an optimiser step tags the bricks it scatters into
        with the step's tag, and :meth:`materialize` visits only those.  ``touched``: a caller-owned flag array (the
        peer-mapped one of ``PeerGradVolume.enable_sparse``), else a local one is allocated."""
        self.get(like)
        self.touched = touched if touched is not None else None
        if self.touched is None:
            self.get_touched(spec)
        self.sparse_sink = True
        self.touch_tag[0] = 1

    def _next_tag(self) -> None:
        self.touch_tag[0] = int(self.touch_tag[0]) % 255 + 1

    def zero(self) -> None:
        if self.buffer is not None and self.dirty:
            self.buffer.zero_()
            if self.sparse_sink:
                self._next_tag()
        self.dirty = False

    def materialize(self, spec: "FusedGridSpec", densities: Tensor, features: Tensor) -> None:
        """``densities.grad`` / ``features.grad`` += what the render backward passes accumulated; then clear."""
        if self.buffer is None or not self.dirty:
            return
        lib = nat.load_library()
        dev = self.buffer.device
        if self.sparse_sink:  # along the trail: adds into .grad and zeroes what it read, so the volume is clean afterwards
            ptrs = []
            for p in (densities, features):
                if p.requires_grad and p.grad is None:
                    p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ptrs.append(p.grad.data_ptr() if p.requires_grad else None)
            if ptrs[0] is not None or ptrs[1] is not None:
                with torch.cuda.device(dev):
                    nat.check(lib.voxe_consume_grad(spec.to_native(), self.buffer.data_ptr(), ptrs[0], ptrs[1], self.touched.data_ptr(),
                                                    int(self.touch_tag[0]), _stream_ptr(dev)), "voxe_consume_grad")
            else:
                self.buffer.zero_()
            self._next_tag()
            self.dirty = False
            return
        outs = []
        for p in (densities, features):
            if not p.requires_grad:
                outs.append((None, 0))
            elif p.grad is None:
                p.grad = torch.empty_like(p, memory_format=torch.contiguous_format)
                outs.append((p.grad, 0))
            else:
                outs.append((p.grad, 1))
        gd = spec.to_native()
        with torch.cuda.device(dev):
            s = _stream_ptr(dev)
            for (g, acc), which in zip(outs, (0, 1)):
                if g is None:
                    continue
                args = (g.data_ptr(), None) if which == 0 else (None, g.data_ptr())
                nat.check(lib.voxe_unpack_grad(gd, self.buffer.data_ptr(), args[0], args[1], acc, s), "voxe_unpack_grad")
        self.buffer.zero_()
        self.dirty = False


def _prep_rays(rays_o: Tensor, rays_d: Tensor) -> Tuple[Tensor, Tensor]:
    assert rays_o.dim() == 2 and rays_d.dim() == 2, "Please note that the RENDER interface only works with FLAT RAYS!"
    assert rays_o.shape == rays_d.shape and rays_o.shape[-1] == 3
    if rays_o.dtype != torch.float32:
        rays_o = rays_o.float()
    if rays_d.dtype != torch.float32:
        rays_d = rays_d.float()
    return rays_o, rays_d


def fused_render(
    gspec: FusedGridSpec,
    rspec: FusedRenderSpec,
    densities: Tensor,
    features: Tensor,
    rays_o: Tensor,
    rays_d: Tensor,
    cache: Optional[PackedVolumeCache] = None,
    jitter: Optional[Tensor] = None,
    noise: Optional[Tensor] = None,
    generator: Optional[torch.Generator] = None,
    grad_sink: Optional[PackedGradAccumulator] = None,
    grad_scratch: Optional[PackedGradAccumulator] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Render flat rays through the fused kernels.  Returns (colour [R,C], depth [R,1], acc [R,1], disparity [R,1]).

    ``jitter`` / ``noise`` default to fresh ``torch.rand`` / ``torch.randn`` draws on the rays' device -- the same
    generator, shapes and order as sample.py:63 and accumulate.py:59-62 -- when the render spec needs them.
    """
    dev = densities.device
    if dev.type != "cuda" or rays_o.device != dev:
        _require_cuda(densities, features, rays_o, rays_d)  # raises the explanatory error
    if densities.dtype != torch.float32 or features.dtype != torch.float32:
        raise TypeError(f"the render path computes in fp32 (got {densities.dtype} / {features.dtype})")
    if rays_o.dtype != torch.float32 or rays_d.dtype != torch.float32:
        rays_o, rays_d = rays_o.float(), rays_d.float()
    # ray / jitter / noise shapes are checked once, in the C++ entry point (same messages as _prep_rays)
    R = rays_o.shape[0]
    ext = _bridge_module or bridge()
    packed = (cache or PackedVolumeCache()).get(gspec, densities, features)
    if R == 0:
        z = torch.zeros((0, 1), dtype=torch.float32, device=dev)
        return torch.zeros((0, rspec.n_colour), dtype=torch.float32, device=dev), z, z.clone(), z.clone()
    volume = flag = touched = tag = None
    mode = ext.MODE_DENSE
    if torch.is_grad_enabled() and (densities.requires_grad or features.requires_grad):
        if grad_sink is not None:  # deferred gradients (opt-in): leave them in the grid's persistent volume
            volume, flag, mode = grad_sink.get(packed), grad_sink.flag, ext.MODE_SINK
            if grad_sink.sparse_sink:  # ... and a trail of the bricks this step touches
                touched, tag = grad_sink.touched, grad_sink.touch_tag
        elif grad_scratch is not None:  # persistent all-zero-between-calls volume: no allocation / zero-fill per call
            volume = grad_scratch.get(packed)
            touched, tag = grad_scratch.touched, grad_scratch.touch_tag
            if touched is None:
                touched = grad_scratch.get_touched(gspec)
            mode = ext.MODE_DIRECT if DIRECT_GRAD_ACCUMULATION else ext.MODE_DENSE
    colour, depth, acc, disparity = ext.render(densities, features, packed, rays_o, rays_d, jitter, noise, volume, flag, touched, tag,
                                               gspec.native_bytes(), rspec.native_bytes(), mode, STRICT_REFERENCE_RNG, generator)
    return colour, depth, acc, disparity


# Early termination of the whole-camera inference render (``fused_render_camera``): a ray stops once its transmittance
# falls below this value; what the skipped samples could still add to a pixel is smaller than it (north-star pixel
# tolerance: 1e-4).  0.0 evaluates every sample.  Never applied to renders that are differentiated.
INFERENCE_MIN_TRANSMITTANCE = 1e-5


def fused_render_camera(
    gspec: FusedGridSpec,
    rspec: FusedRenderSpec,
    densities: Tensor,
    features: Tensor,
    height: int,
    width: int,
    focal: float,
    rotation,
    translation,
    cache: Optional[PackedVolumeCache] = None,
    first_pixel: int = 0,
    num_pixels: Optional[int] = None,
    generator: Optional[torch.Generator] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Forward-only render of a pinhole camera through ``voxe_render_camera``: rays are generated inside the kernel
    (``cast_rays``, misc.py:12-50), one launch for the pixel range.  Returns flat (colour [P,C], depth [P,1], acc [P,1],
    disparity [P,1]) in ``flatten_rays`` order.  ``rotation`` [3,3] / ``translation`` [3] or [3,1]: tensors or arrays."""
    dev = densities.device
    if dev.type != "cuda":
        _require_cuda(densities, features)
    if rspec.noise_std != 0.0:
        raise NotImplementedError("fused_render_camera does not take density noise; render rays through fused_render")
    lib = nat.load_library()
    packed = (cache or PackedVolumeCache()).get(gspec, densities, features)
    total = int(height) * int(width)
    num_pixels = total - first_pixel if num_pixels is None else int(num_pixels)
    rot = np.asarray(rotation.detach().cpu() if isinstance(rotation, Tensor) else rotation, dtype=np.float32).reshape(3, 3)
    trans = np.asarray(translation.detach().cpu() if isinstance(translation, Tensor) else translation, dtype=np.float32).reshape(3)
    cam = nat.VoxeCameraDesc()
    cam.height, cam.width, cam.focal = int(height), int(width), float(focal)
    cam.rotation[:] = rot.reshape(-1).tolist()
    cam.translation[:] = trans.tolist()
    rd = nat.VoxeRenderDesc.from_buffer_copy(rspec.native_bytes())
    if rspec.flags & nat.FLAG_PERTURB:  # in-kernel draws: (seed, offset) from the device generator, which is advanced
        gen = generator if generator is not None else torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        rd.rng_seed, rd.rng_offset = int(gen.initial_seed()) & (2**64 - 1), int(gen.get_offset())
        gen.set_offset(gen.get_offset() + 4)
    colour = torch.empty((num_pixels, rspec.n_colour), dtype=torch.float32, device=dev)
    depth = torch.empty((num_pixels, 1), dtype=torch.float32, device=dev)
    acc = torch.empty((num_pixels, 1), dtype=torch.float32, device=dev)
    disp = torch.empty((num_pixels, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nat.check(
            lib.voxe_render_camera(gspec.to_native(), rd, cam, packed.data_ptr(), int(first_pixel), num_pixels, colour.data_ptr(),
                                   depth.data_ptr(), acc.data_ptr(), disp.data_ptr(), float(INFERENCE_MIN_TRANSMITTANCE), _stream_ptr(dev)),
            "voxe_render_camera",
        )
    return colour, depth, acc, disp


def fused_render_infer(gspec: FusedGridSpec, rspec: FusedRenderSpec, densities: Tensor, features: Tensor, rays_o: Tensor,
                       rays_d: Tensor, cache: Optional[PackedVolumeCache] = None, jitter: Optional[Tensor] = None,
                       generator: Optional[torch.Generator] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Forward-only render of caller-supplied flat rays through the inference kernel (``voxe_render_infer``: one thread
    per ray, early termination at ``INFERENCE_MIN_TRANSMITTANCE``).  Same outputs as ``fused_render``; nothing to
    differentiate."""
    dev = densities.device
    if dev.type != "cuda" or rays_o.device != dev:
        _require_cuda(densities, features, rays_o, rays_d)
    if rspec.noise_std != 0.0:
        raise NotImplementedError("fused_render_infer does not take density noise; render rays through fused_render")
    rays_o, rays_d = _prep_rays(rays_o, rays_d)
    rays_o, rays_d = rays_o.detach().contiguous(), rays_d.detach().contiguous()
    lib = nat.load_library()
    packed = (cache or PackedVolumeCache()).get(gspec, densities, features)
    R = rays_o.shape[0]
    rd = nat.VoxeRenderDesc.from_buffer_copy(rspec.native_bytes())
    if rspec.flags & nat.FLAG_PERTURB:
        if jitter is not None:
            assert jitter.shape == (R, rspec.num_samples)
            jitter = jitter.detach().float().contiguous()
        else:
            gen = generator if generator is not None else torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
            rd.rng_seed, rd.rng_offset = int(gen.initial_seed()) & (2**64 - 1), int(gen.get_offset())
            gen.set_offset(gen.get_offset() + 4)
    else:
        jitter = None
    colour = torch.empty((R, rspec.n_colour), dtype=torch.float32, device=dev)
    depth = torch.empty((R, 1), dtype=torch.float32, device=dev)
    acc = torch.empty((R, 1), dtype=torch.float32, device=dev)
    disp = torch.empty((R, 1), dtype=torch.float32, device=dev)
    if R:
        with torch.cuda.device(dev):
            nat.check(
                lib.voxe_render_infer(gspec.to_native(), rd, packed.data_ptr(), rays_o.data_ptr(), rays_d.data_ptr(), _ptr(jitter),
                                      colour.data_ptr(), depth.data_ptr(), acc.data_ptr(), disp.data_ptr(), R,
                                      float(INFERENCE_MIN_TRANSMITTANCE), _stream_ptr(dev)),
                "voxe_render_infer",
            )
    return colour, depth, acc, disp


def fused_render_attn(*args, **kwargs):
    """Attention-grid twin (renderers.py:108-163): same kernels with VOXE_FLAG_ATTN set in the render spec and the
    1-channel attention volume passed as ``features``."""
    return fused_render(*args, **kwargs)


class _PointQuery(torch.autograd.Function):
    """``voxe_query_points`` / ``voxe_query_points_bwd`` under autograd (differentiable w.r.t. the two grid tensors)."""

    @staticmethod
    def forward(ctx, densities: Tensor, features: Tensor, points: Tensor, gspec: FusedGridSpec, packed: Tensor) -> Tensor:
        dev = densities.device
        lib = nat.load_library()
        n = points.shape[0]
        out = torch.empty((n, gspec.n_features + 1), dtype=torch.float32, device=dev)
        if n:
            with torch.cuda.device(dev):
                nat.check(lib.voxe_query_points(gspec.to_native(), packed.data_ptr(), points.data_ptr(), out.data_ptr(), n, _stream_ptr(dev)),
                          "voxe_query_points")
        # saved (not just referenced): a live graph must keep PackedVolumeCache from repacking this buffer in place
        ctx.save_for_backward(points, packed)
        ctx.gspec, ctx.shapes = gspec, (densities.shape, features.shape)
        return out

    @staticmethod
    def backward(ctx, g_out: Tensor):
        points, packed = ctx.saved_tensors
        gspec, (dshape, fshape) = ctx.gspec, ctx.shapes
        dev = packed.device
        lib = nat.load_library()
        need_d, need_f = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_d = torch.empty(dshape, dtype=torch.float32, device=dev) if need_d else None
        g_f = torch.empty(fshape, dtype=torch.float32, device=dev) if need_f else None
        volume = torch.zeros_like(packed)
        gd = gspec.to_native()
        with torch.cuda.device(dev):
            s = _stream_ptr(dev)
            if points.shape[0]:
                nat.check(lib.voxe_query_points_bwd(gd, packed.data_ptr(), points.data_ptr(), g_out.float().contiguous().data_ptr(),
                                                    volume.data_ptr(), points.shape[0], s), "voxe_query_points_bwd")
            nat.check(lib.voxe_unpack_grad(gd, volume.data_ptr(), _ptr(g_d), _ptr(g_f), 0, s), "voxe_unpack_grad")
        return g_d, g_f, None, None, None


def query_points(gspec: FusedGridSpec, densities: Tensor, features: Tensor, points: Tensor,
                 cache: Optional[PackedVolumeCache] = None) -> Tensor:
    """Stand-alone point query (voxels.py:287-345 upstream): ``[N, F + 1]`` = interpolated features and post-activated
    interpolated density at ``points`` [N, 3] (world coordinates, anywhere: zeros padding, no inside mask)."""
    assert points.dim() == 2 and points.shape[-1] == 3, f"points should be of shape [N x 3] as opposed to ({tuple(points.shape)})"
    _require_cuda(densities, features)
    if points.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("voxe_query_points differentiates the grid tensors only; detach the points (no caller of the "
                                  "reference differentiates a query w.r.t. its coordinates)")
    if points.device != densities.device:
        raise RuntimeError(f"all render inputs must live on one device (got {points.device} and {densities.device})")
    packed = (cache or PackedVolumeCache()).get(gspec, densities, features)
    return _PointQuery.apply(densities, features, points.detach().float().contiguous(), gspec, packed)
