"""voxe_b200 -- Python binding of libvoxe_sm100a.so, the B200-native ray-marcher behind Vox-E's render API.

Layout of the product tree (``vox-e_b200/`` is a path root, put it on ``sys.path``):

    csrc/          CUDA kernels + the C ABI declared in ``include/voxe.h``
    voxe_b200/     this package: ctypes loader (_native), autograd bridge (render_function), fused optimiser step (optim),
                   grid regularisers (regularizers), ray-batch sampler (sampling), ray-shard data parallelism (dist)
    thre3d_atom/   the reference-facing interface for the hot path (same module paths and names as Vox-E, so its
                   scripts, identity asserts and pickled checkpoints resolve to the fused implementation)

There is no CPU implementation in this tree: rendering needs the compiled library and a CUDA device, and says so.
"""
from voxe_b200._native import NativeLibraryError, library_path, load_library  # noqa: F401
from voxe_b200.render_function import FusedGridSpec, fused_render, fused_render_attn  # noqa: F401

__all__ = ["NativeLibraryError", "library_path", "load_library", "FusedGridSpec", "fused_render", "fused_render_attn"]
