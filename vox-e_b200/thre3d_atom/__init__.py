"""Reference-facing interface of the fused ray-marcher.

This package deliberately answers to the import paths of Vox-E's own ``thre3d_atom`` for everything on the render
hot path: the reference's training / editing / rendering scripts, their identity asserts on
``render_sh_voxel_grid`` (modules/trainers.py:127-129) and checkpoints that pickle those import paths resolve here,
and land in ``voxe_b200`` (CUDA, sm_100a) instead of the stock-ATen pipeline.  Only the hot path and its boundary are
provided; trainers, diffusion guidance, datasets and visualisation stay with the reference.
"""
