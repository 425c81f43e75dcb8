"""ctypes loader for ``libvoxe_sm100a.so`` (C ABI: ``include/voxe.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C vox-e_b200/csrc`` and is the only compute
backend: if it is missing or fails to load, :func:`load_library` raises -- nothing falls back to PyTorch or the CPU.
"""
from __future__ import annotations

import ctypes
import os
import threading
from pathlib import Path
from typing import Optional

_LIB_NAME = "libvoxe_sm100a.so"
_lock = threading.Lock()
_lib: Optional[ctypes.CDLL] = None

# ---- enums of include/voxe.h -------------------------------------------------------------------------------
ABI_VERSION = 14
PREACT_IDENTITY, PREACT_ABS = 0, 1
POSTACT_IDENTITY, POSTACT_RELU, POSTACT_SOFTPLUS = 0, 1, 2
FLAG_PERTURB, FLAG_AABB_SAMPLING, FLAG_DISPARITY_SAMPLING = 1, 2, 4
FLAG_WHITE_BKGD, FLAG_RENDER_DIFFUSE, FLAG_ATTN = 8, 16, 32
PAIR_CORRELATION, PAIR_L2, PAIR_L1 = 0, 1, 2
REG_WORKSPACE_DOUBLES = 8192


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing, stale, or a call into it failed."""


class VoxeGridDesc(ctypes.Structure):
    _fields_ = [
        ("dims", ctypes.c_int32 * 3),
        ("n_features", ctypes.c_int32),
        ("channels", ctypes.c_int32),
        ("aabb_lo", ctypes.c_float * 3),
        ("aabb_hi", ctypes.c_float * 3),
        ("norm_scale", ctypes.c_float * 3),
        ("norm_bias", ctypes.c_float * 3),
        ("density_scale", ctypes.c_float),
        ("preact", ctypes.c_int32),
        ("postact", ctypes.c_int32),
    ]


class VoxeRenderDesc(ctypes.Structure):
    _fields_ = [
        ("num_samples", ctypes.c_int32),
        ("near", ctypes.c_float),
        ("far", ctypes.c_float),
        ("flags", ctypes.c_int32),
        ("sh_degree", ctypes.c_int32),
        ("n_colour", ctypes.c_int32),
        ("noise_std", ctypes.c_float),
        ("rng_seed", ctypes.c_uint64),
        ("rng_offset", ctypes.c_uint64),
        ("rng_seed_dev", ctypes.c_void_p),
        ("rng_offset_dev", ctypes.c_void_p),
        ("rng_offset_intragraph", ctypes.c_uint64),
        ("stats", ctypes.c_void_p),
    ]


class VoxeCameraDesc(ctypes.Structure):
    _fields_ = [("height", ctypes.c_int32), ("width", ctypes.c_int32), ("focal", ctypes.c_float), ("rotation", ctypes.c_float * 9),
                ("translation", ctypes.c_float * 3)]


class VoxeSamplerDesc(ctypes.Structure):
    _fields_ = [("num_pixels", ctypes.c_int64), ("height", ctypes.c_int32), ("width", ctypes.c_int32), ("focal", ctypes.c_float),
                ("pixel_channels", ctypes.c_int32), ("rng_seed", ctypes.c_uint64), ("rng_offset", ctypes.c_uint64)]


MAX_PEERS, SIGNAL_WORDS, NCCL_UNIQUE_ID_BYTES = 16, 4096, 128


class VoxePeerDesc(ctypes.Structure):
    _fields_ = [("world_size", ctypes.c_int32), ("rank", ctypes.c_int32), ("buffers", ctypes.c_void_p * MAX_PEERS),
                ("signals", ctypes.c_void_p * MAX_PEERS), ("multicast", ctypes.c_void_p), ("multicast_share", ctypes.c_int32)]


class VoxeAdamDesc(ctypes.Structure):
    _fields_ = [("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double), ("eps", ctypes.c_double),
                ("step", ctypes.c_int32)]


# every symbol include/voxe.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_GD, _RD = ctypes.POINTER(VoxeGridDesc), ctypes.POINTER(VoxeRenderDesc)
EXPORTS = {
    "voxe_abi_version": (ctypes.c_int, []),
    "voxe_last_error": (ctypes.c_char_p, []),
    "voxe_packed_channels": (ctypes.c_int, [ctypes.c_int]),
    "voxe_packed_floats": (ctypes.c_int64, [_GD]),
    "voxe_pack_grid": (ctypes.c_int, [_GD, _P, _P, _P, _P]),
    "voxe_unpack_grad": (ctypes.c_int, [_GD, _P, _P, _P, ctypes.c_int, _P]),
    "voxe_consume_grad": (ctypes.c_int, [_GD, _P, _P, _P, _P, ctypes.c_int32, _P]),
    "voxe_touched_bytes": (ctypes.c_int64, [_GD]),
    "voxe_jitter_fill": (ctypes.c_int, [_RD, _P, ctypes.c_int64, _P]),
    "voxe_adam_step": (ctypes.c_int, [_GD, ctypes.POINTER(VoxeAdamDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "voxe_saved_floats": (ctypes.c_int64, [_RD, ctypes.c_int64]),
    "voxe_render_fwd": (ctypes.c_int, [_GD, _RD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, _P]),
    "voxe_render_camera": (ctypes.c_int, [_GD, _RD, ctypes.POINTER(VoxeCameraDesc), _P, ctypes.c_int64, ctypes.c_int64, _P, _P, _P, _P,
                                          ctypes.c_float, _P]),
    "voxe_render_infer": (ctypes.c_int, [_GD, _RD, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int64, ctypes.c_float, _P]),
    "voxe_render_bwd": (ctypes.c_int, [_GD, _RD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int32, ctypes.c_int64, _P]),
    "voxe_resample_grid": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int32 * 3), ctypes.c_int32, _P, ctypes.POINTER(ctypes.c_int32 * 3), _P]),
    "voxe_query_points": (ctypes.c_int, [ctypes.POINTER(VoxeGridDesc), _P, _P, _P, ctypes.c_int64, _P]),
    "voxe_query_points_bwd": (ctypes.c_int, [ctypes.POINTER(VoxeGridDesc), _P, _P, _P, _P, ctypes.c_int64, _P]),
    "voxe_tv_regularizer": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int32 * 3), ctypes.c_int32, ctypes.c_int32, _P, _P, _P,
                                           ctypes.c_float, _P, ctypes.c_int32, _P]),
    "voxe_pair_loss": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int32, _P, _P, _P, _P]),
    "voxe_pair_loss_grad": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int32, _P, _P, ctypes.c_float, _P, ctypes.c_int32, _P]),
    "voxe_sample_rays": (ctypes.c_int, [ctypes.POINTER(VoxeSamplerDesc), _P, _P, _P, _P, _P, ctypes.c_int64, _P, _P, _P, _P, _P]),
    "voxe_allreduce_grads_peer": (ctypes.c_int, [ctypes.POINTER(VoxePeerDesc), ctypes.c_int64, _P, _P]),
    "voxe_peer_touched_bytes": (ctypes.c_int64, [_GD]),
    "voxe_allreduce_grads_peer_sparse": (ctypes.c_int, [ctypes.POINTER(VoxePeerDesc), _GD, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, _P, _P]),
    "voxe_nccl_unique_id": (ctypes.c_int, [_P]),
    "voxe_nccl_comm_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32, _P]),
    "voxe_nccl_comm_destroy": (ctypes.c_int, [_P]),
    "voxe_allreduce_grads": (ctypes.c_int, [_P, _P, ctypes.c_int64, _P]),
    "voxe_set_tuning": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "voxe_launch_count": (ctypes.c_int64, []),
    "voxe_specialised_launch_count": (ctypes.c_int64, []),
}


def library_path() -> Path:
    override = os.environ.get("VOXE_LIBRARY")
    return Path(override) if override else Path(__file__).resolve().parent / _LIB_NAME


def load_library() -> ctypes.CDLL:
    """Load (once) and type the C ABI.  Raises NativeLibraryError when the .so is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not path.exists():
            raise NativeLibraryError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                f"`make -C vox-e_b200/csrc`.  There is no CPU / PyTorch fallback for the render path."
            )
        try:
            lib = ctypes.CDLL(str(path))
        except OSError as exc:  # e.g. libcudart missing
            raise NativeLibraryError(f"cannot load {path}: {exc}") from exc
        for name, (restype, argtypes) in EXPORTS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{path} does not export {name}; rebuild the library") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.voxe_abi_version() != ABI_VERSION:
            raise NativeLibraryError(f"{path} has ABI {lib.voxe_abi_version()}, binding expects {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load_library().voxe_last_error().decode(errors="replace")
        if code == 2:  # VOXE_ERR_UNSUPPORTED
            raise NotImplementedError(f"{what}: {msg}")
        raise NativeLibraryError(f"{what} failed with code {code}: {msg}")


def launch_count() -> int:
    return int(load_library().voxe_launch_count())


def set_tuning(samples_per_thread: int = 0, rays_per_cta: int = 0, register_cap: int = 0) -> None:
    """Launch-shape override for tuning runs (0 = built-in choice).  Do not change it between a forward call and its
    backward: the ``saved`` workspace layout depends on it."""
    check(load_library().voxe_set_tuning(samples_per_thread, rays_per_cta, register_cap), "voxe_set_tuning")
