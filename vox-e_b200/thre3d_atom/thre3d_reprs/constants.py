"""Names used inside saved-model dictionaries.

Checkpoints written by Vox-E and by this package must load in each other (``VolumetricModel.get_save_info`` /
``create_volumetric_model_from_saved_model``), so the key strings are an on-disk contract shared with the reference
(thre3d_atom/thre3d_reprs/constants.py:1-16); they are generated here from two naming rules instead of being listed:

* top-level dictionary keys are the lower-cased constant names (``THRE3D_REPR -> "thre3d_repr"``);
* ``u_<NAME>`` constants name entries of a representation's ``state_dict``: tensors that are private attributes carry a
  leading underscore (``u_DENSITIES -> "_densities"``), sub-modules do not (``u_ATTN -> "attn"``).
"""

_TOP_LEVEL = ("THRE3D_REPR", "RENDER_PROCEDURE", "RENDER_CONFIG", "RENDER_CONFIG_TYPE", "STATE_DICT", "CONFIG_DICT")
_PRIVATE_TENSORS = ("DENSITIES", "FEATURES", "IN_DENSITIES", "IN_FEATURES")
_SUBMODULES = ("RGBNET", "DENSITYNET", "ATTN")

globals().update({name: name.lower() for name in _TOP_LEVEL})
globals().update({f"u_{name}": f"_{name.lower()}" for name in _PRIVATE_TENSORS})
globals().update({f"u_{name}": name.lower() for name in _SUBMODULES})

__all__ = [*_TOP_LEVEL, *(f"u_{name}" for name in _PRIVATE_TENSORS + _SUBMODULES)]
