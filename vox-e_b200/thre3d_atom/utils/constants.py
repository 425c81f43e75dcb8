"""Constants of the render path.  Names and values are part of the interface shared with the reference
(thre3d_atom/utils/constants.py:1-28): the ``extra`` keys of a ``RenderOut`` and the ``extra_info`` keys of a checkpoint are
read by the reference's trainers, testers and visualisers, and ``ZERO_PLUS`` / ``INFINITY`` are baked into the CUDA kernels
(``kZeroPlus`` / ``kInfinity`` in csrc/voxe_device.cuh)."""

# channel counts -----------------------------------------------------------------------------------------------------
NUM_COORD_DIMENSIONS, NUM_COLOUR_CHANNELS, NUM_RGBA_CHANNELS, NUM_ATTN_CHANNELS = 3, 3, 4, 1

# numerics -----------------------------------------------------------------------------------------------------------
SEED = 42                # conftest / script seeding
ZERO_PLUS = 1e-10        # added to denominators (ray slab test, disparity)
INFINITY = 1e10          # length of the last sample interval; colour logit of masked samples is -INFINITY

# RenderOut.extra: the two maps every render returns, then the per-point debug maps of ``extra_debug_info`` -----------------
_EXTRA = dict(
    EXTRA_DISPARITY="disparity",
    EXTRA_ACCUMULATED_WEIGHTS="accumulated_weight",
    EXTRA_POINT_DENSITIES="point_densities",
    EXTRA_POINT_OCCUPANCIES="point_occupancies",
    EXTRA_SAMPLE_INTERVALS="deltas",
    EXTRA_POINT_WEIGHTS="point_weights",
    EXTRA_POINT_DEPTHS="point_depths",
)
# checkpoint ``extra_info`` and dataset metadata: the key is the lower-cased name ------------------------------------------
_LOWER_CASED = ("CAMERA_BOUNDS", "CAMERA_INTRINSICS", "HEMISPHERICAL_RADIUS", "EXTRA_INFO")

globals().update(_EXTRA)
globals().update({name: name.lower() for name in _LOWER_CASED})
