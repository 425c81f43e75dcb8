"""VolumetricModel: the facade every caller of the render path goes through
(reference: thre3d_atom/modules/volumetric_model.py:30-301).

``render_rays`` is the differentiable entry (training / SDS editing); ``render`` casts a whole camera and walks it in
ray chunks under ``torch.no_grad()``.  Overrides arrive as keyword arguments and are validated against the config's
attributes (``ValueError`` for unknown names).  The saved-model dictionary layout is unchanged, so checkpoints written by
either implementation load in the other (``torch.load(..., weights_only=False)`` is needed on torch >= 2.6 because the
payload pickles functions and NamedTuples).
"""
import copy
import dataclasses
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Tuple

import torch
from torch.nn import Module

from thre3d_atom.rendering.volumetric.render_interface import Rays, RenderOut, RenderOutAttn
from thre3d_atom.rendering.volumetric.utils.misc import (
    cast_rays,
    collate_rendered_output,
    collate_rendered_output_attn,
    flatten_rays,
    reshape_rendered_output,
    reshape_rendered_output_attn,
)
from thre3d_atom.thre3d_reprs.constants import (
    CONFIG_DICT,
    RENDER_CONFIG,
    RENDER_CONFIG_TYPE,
    RENDER_PROCEDURE,
    STATE_DICT,
    THRE3D_REPR,
)
from thre3d_atom.thre3d_reprs.renderers import RenderConfig, RenderProcedure, render_sh_voxel_grid_attn, render_sh_voxel_grid_camera
from thre3d_atom.utils.constants import EXTRA_INFO
from thre3d_atom.utils.imaging_utils import CameraIntrinsics, CameraPose


class VolumetricModel:
    def __init__(
        self,
        thre3d_repr: Module,
        render_procedure: RenderProcedure,
        render_config: RenderConfig,
        render_procedure_attn=None,
        device: torch.device = torch.device("cuda" if torch.cuda.is_available() else "cpu"),
    ) -> None:
        self._thre3d_repr = thre3d_repr.to(device)
        self._render_procedure = render_procedure
        self._render_procedure_attn = render_procedure_attn
        self._render_config = render_config
        self._device = device

    @property
    def thre3d_repr(self) -> Module:
        return self._thre3d_repr

    @thre3d_repr.setter
    def thre3d_repr(self, thre3d_repr: Module) -> None:
        self._thre3d_repr = thre3d_repr

    @property
    def render_procedure(self) -> RenderProcedure:
        return self._render_procedure

    @property
    def render_config(self) -> RenderConfig:
        return self._render_config

    @property
    def device(self) -> torch.device:
        return self._device

    @staticmethod
    def _update_render_config(render_config: RenderConfig, update_dict: Dict[str, Any]) -> RenderConfig:
        """A private copy of the config with ``update_dict`` applied; the stored config is never touched."""
        if not update_dict:
            return render_config  # render procedures never mutate the config they are handed
        updated = copy.deepcopy(render_config)
        for field, value in update_dict.items():
            if not hasattr(updated, field):
                raise ValueError(f"Unknown render configuration field {field} requested for overriding :(")
            setattr(updated, field, value)
        return updated

    def get_save_info(self, extra_info: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        save_info = {
            THRE3D_REPR: {
                STATE_DICT: self._thre3d_repr.state_dict(),
                CONFIG_DICT: self._thre3d_repr.get_save_config_dict(),
            },
            RENDER_PROCEDURE: self._render_procedure,
            RENDER_CONFIG_TYPE: type(self._render_config),
            RENDER_CONFIG: dataclasses.asdict(self._render_config),
        }
        if extra_info is not None:
            save_info[EXTRA_INFO] = extra_info
        return save_info

    # ------------------------------------------------------------------------------------------------------
    # differentiable ray renders
    # ------------------------------------------------------------------------------------------------------
    def render_rays(self, rays: Rays, parallel_points_chunk_size: Optional[int] = None, **kwargs) -> RenderOut:
        render_config = self._update_render_config(self._render_config, kwargs)
        return self._render_procedure(self._thre3d_repr, rays, render_config, parallel_points_chunk_size)

    def render_rays_attn(
        self, rays: Rays, parallel_points_chunk_size: Optional[int] = None, orig_densities=False, **kwargs
    ) -> RenderOutAttn:
        render_config = self._update_render_config(self._render_config, kwargs)
        return self._render_procedure_attn(self._thre3d_repr, rays, render_config, parallel_points_chunk_size, orig_densities)

    # ------------------------------------------------------------------------------------------------------
    # whole-camera renders (no grad)
    # ------------------------------------------------------------------------------------------------------
    def _whole_camera_in_one_launch(self, noise_std: float, attn: bool) -> bool:
        """True for the fused render procedures, unless a call would have to materialise [R, S] random draws
        (reference-RNG replay, density noise); any other procedure keeps the caller's chunking."""
        from thre3d_atom.thre3d_reprs.renderers import render_sh_voxel_grid
        from voxe_b200 import render_function as rf

        procedure = self._render_procedure_attn if attn else self._render_procedure
        fused = procedure is (render_sh_voxel_grid_attn if attn else render_sh_voxel_grid)
        return fused and not rf.STRICT_REFERENCE_RNG and noise_std == 0.0

    def _render_camera(self, render_chunk, collate, reshape, camera_pose, camera_intrinsics, chunk_size, gpu_render, verbose,
                       kwargs_noise_std=0.0, attn=False, camera_fast_path=None):
        if camera_fast_path is not None and self._whole_camera_in_one_launch(kwargs_noise_std, attn):
            # fused procedure, nothing to differentiate: rays are generated inside the kernel and the camera is one launch
            with torch.no_grad():
                out = camera_fast_path()
            return reshape(out if gpu_render else out.to(torch.device("cpu")), camera_intrinsics=camera_intrinsics)
        flat_rays = flatten_rays(cast_rays(camera_intrinsics=camera_intrinsics, pose=camera_pose, device=self._device))
        chunk_size = len(flat_rays) if chunk_size is None else chunk_size
        if self._whole_camera_in_one_launch(kwargs_noise_std, attn):
            # The reference chunks rays (32768 per call, volumetric_model.py:170-186) to bound the [R, S, .] tensors of its
            # pipeline; the fused forward kernel keeps O(R) state, so under no_grad the whole camera is ONE launch and,
            # with gpu_render=False, one device->host copy instead of one per chunk.
            chunk_size = max(chunk_size, len(flat_rays))
        starts = range(0, len(flat_rays), chunk_size)
        if verbose:
            from tqdm import tqdm

            starts = tqdm(starts)
        chunks = []
        with torch.no_grad():
            for start in starts:
                chunk = render_chunk(flat_rays[start : start + chunk_size])
                chunks.append(chunk if gpu_render else chunk.to(torch.device("cpu")))
        return reshape(collate(chunks), camera_intrinsics=camera_intrinsics)

    def render(
        self,
        camera_pose: CameraPose,
        camera_intrinsics: CameraIntrinsics,
        parallel_rays_chunk_size: Optional[int] = 32768,
        parallel_points_chunk_size: Optional[int] = None,
        gpu_render: bool = True,
        verbose: bool = False,
        **kwargs,
    ) -> RenderOut:
        """[H, W, .] colour / depth / extras of one camera; ``kwargs`` override render-config fields for this call."""
        return self._render_camera(
            lambda rays: self.render_rays(rays, parallel_points_chunk_size, **kwargs),
            collate_rendered_output, reshape_rendered_output,
            camera_pose, camera_intrinsics, parallel_rays_chunk_size, gpu_render, verbose,
            kwargs_noise_std=float(kwargs.get("stochastic_density_noise_std", getattr(self._render_config, "stochastic_density_noise_std", 0.0))),
            camera_fast_path=lambda: render_sh_voxel_grid_camera(
                self._thre3d_repr, camera_intrinsics, camera_pose, self._update_render_config(self._render_config, kwargs)),
        )

    def render_attn(
        self,
        camera_pose: CameraPose,
        camera_intrinsics: CameraIntrinsics,
        parallel_rays_chunk_size: Optional[int] = 32768,
        parallel_points_chunk_size: Optional[int] = None,
        gpu_render: bool = True,
        verbose: bool = False,
        orig_densities=False,
        **kwargs,
    ) -> RenderOutAttn:
        return self._render_camera(
            lambda rays: self.render_rays_attn(rays, parallel_points_chunk_size, orig_densities, **kwargs),
            collate_rendered_output_attn, reshape_rendered_output_attn,
            camera_pose, camera_intrinsics, parallel_rays_chunk_size, gpu_render, verbose,
            kwargs_noise_std=float(kwargs.get("stochastic_density_noise_std", getattr(self._render_config, "stochastic_density_noise_std", 0.0))),
            attn=True,
        )


def _load_saved_model(model_path: Path, make_repr: Callable[[Dict[str, Any]], Module], device: torch.device):
    model_data = torch.load(model_path, weights_only=False)  # pickled functions / NamedTuples inside
    render_config = model_data[RENDER_CONFIG_TYPE](**model_data[RENDER_CONFIG])
    vol_mod = VolumetricModel(
        thre3d_repr=make_repr(model_data),
        render_procedure=model_data[RENDER_PROCEDURE],
        render_procedure_attn=render_sh_voxel_grid_attn,
        render_config=render_config,
        device=device,
    )
    return vol_mod, model_data[EXTRA_INFO]


def create_volumetric_model_from_saved_model(
    model_path: Path,
    thre3d_repr_creator: Callable[[Dict[str, Any]], Module],
    device: torch.device = torch.device("cpu"),
) -> Tuple[VolumetricModel, Dict[str, Any]]:
    return _load_saved_model(model_path, thre3d_repr_creator, device)


def create_volumetric_model_from_saved_model_attn(
    model_path: Path,
    thre3d_repr_creator: Callable[[Dict[str, Any]], Module],
    device: torch.device = torch.device("cpu"),
    load_attn=False,
) -> Tuple[VolumetricModel, Dict[str, Any]]:
    return _load_saved_model(model_path, lambda data: thre3d_repr_creator(data, load_attn=load_attn), device)
