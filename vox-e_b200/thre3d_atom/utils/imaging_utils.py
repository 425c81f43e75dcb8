"""Camera model helpers used by the render path and its harnesses.

Reference: thre3d_atom/utils/imaging_utils.py -- camera tuples :17-30, ``adjust_dynamic_range`` :42-71, intrinsics scaling
:140-150, pose construction :153-232.  ``postprocess_depth_map`` (:93-125, depth colouring for the reference's
visualisation modules) is kept so that ``visualizations/{static,animations}.py`` import against this module; matplotlib
is imported when it is called, not at module import.
"""
import math
from typing import NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from thre3d_atom.utils.constants import NUM_COLOUR_CHANNELS


class CameraIntrinsics(NamedTuple):
    height: int
    width: int
    focal: float


class CameraPose(NamedTuple):
    rotation: np.array  # [3, 3]
    translation: np.array  # [3, 1]


class CameraBounds(NamedTuple):
    near: float
    far: float


def to8b(x: np.array) -> np.array:
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def adjust_dynamic_range(
    data: Union[np.array, Tensor],
    drange_in: Tuple[float, float],
    drange_out: Tuple[float, float],
    slack: bool = False,
) -> Union[np.array, Tensor]:
    """Affine map of ``data`` from ``drange_in`` onto ``drange_out``.  With ``slack`` the map is one fp32 scale and bias
    (values outside the input range extrapolate); without it the result is clipped to the output range."""
    if drange_in == drange_out:
        return data
    in_lo, in_hi = np.float32(drange_in[0]), np.float32(drange_in[1])
    out_lo, out_hi = np.float32(drange_out[0]), np.float32(drange_out[1])
    if slack:
        scale = (out_hi - out_lo) / (in_hi - in_lo)
        bias = out_lo - in_lo * scale
        return data * scale + bias
    mapped = ((data - in_lo) / (in_hi - in_lo) * (out_hi - out_lo)) + out_lo
    return mapped.clip(drange_out[0], drange_out[1])


def get_2d_coordinates(height: int, width: int, drange: Tuple[float, float] = (-1.0, 1.0)) -> Tensor:
    lo, hi = drange
    rows = torch.linspace(lo, hi, height, dtype=torch.float32)
    cols = torch.linspace(lo, hi, width, dtype=torch.float32)
    return torch.stack(torch.meshgrid(rows, cols, indexing="ij"), dim=-1)


def postprocess_depth_map(
    depth_map: np.array, camera_bounds: Optional[CameraBounds] = None, acc_map: Optional[np.array] = None
) -> np.array:
    """Colour a depth map with the "magma" colour map (8-bit RGB).  With ``acc_map`` the range tops out at the largest
    *foreground* depth and the result is composited over white with the squared-transparency weighting of
    imaging_utils.py:117-122.  ``camera_bounds`` is accepted and unused, as upstream."""
    import matplotlib.pyplot as plt  # visualisation-only dependency, deliberately not a module-level import

    if depth_map.ndim == 3 and depth_map.shape[-1] == 1:
        depth_map = depth_map[..., 0]
    lo = depth_map.min()
    hi = (depth_map * acc_map[..., 0]).max() if acc_map is not None else depth_map.max()
    unit = adjust_dynamic_range(depth_map, drange_in=(lo, hi), drange_out=(0, 1), slack=True)
    coloured = plt.get_cmap("magma", lut=1024)(unit)[..., :NUM_COLOUR_CHANNELS]
    if acc_map is None:
        return to8b(coloured)
    see_through = (1.0 - acc_map) ** 2
    return to8b((coloured * acc_map + see_through) / (acc_map + see_through))


def scale_camera_intrinsics(camera_intrinsics: CameraIntrinsics, scale_factor: float = 1.0) -> CameraIntrinsics:
    return CameraIntrinsics(
        height=int(np.ceil(camera_intrinsics.height * scale_factor)),
        width=int(np.ceil(camera_intrinsics.width * scale_factor)),
        focal=camera_intrinsics.focal * scale_factor,
    )


def _homogeneous(rows, device) -> Tensor:
    return torch.tensor(rows, dtype=torch.float32, device=device)


def _translate_z(z: float, device=torch.device("cpu")) -> Tensor:
    return _homogeneous([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, z], [0.0, 0.0, 0.0, 1.0]], device)


def _rotate_pitch(pitch: float, device=torch.device("cpu")) -> Tensor:
    c, s = np.cos(pitch), np.sin(pitch)
    return _homogeneous([[1.0, 0.0, 0.0, 0.0], [0.0, c, -s, 0.0], [0.0, s, c, 0.0], [0.0, 0.0, 0.0, 1.0]], device)


def _rotate_yaw(yaw: float, device=torch.device("cpu")) -> Tensor:
    c, s = np.cos(yaw), np.sin(yaw)
    return _homogeneous([[c, -s, 0.0, 0.0], [s, c, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]], device)


def _camera_to_world(yaw_deg: float, pitch_deg: float, radius: float, device) -> Tensor:
    c2w = _translate_z(radius, device)
    c2w = _rotate_pitch(pitch_deg / 180.0 * np.pi, device) @ c2w
    return _rotate_yaw(yaw_deg / 180.0 * np.pi, device) @ c2w


def pose_spherical(yaw: float, pitch: float, radius: float, device=torch.device("cpu")) -> CameraPose:
    """Camera on a sphere of ``radius`` looking at the origin; angles in degrees."""
    c2w = _camera_to_world(yaw, pitch, radius, device)
    return CameraPose(rotation=c2w[:3, :3], translation=c2w[:3, 3:])


def get_random_pose(radius: float, device=torch.device("cpu")):
    """Random view for score-distillation editing: pitch U(15, 90), yaw U(0, 360) drawn from numpy's global RNG in that
    order.  Returns (pose, view-dependent prompt word, pitch, yaw) like the reference."""
    rand_pitch = 15.0 + float(np.random.rand(1)[0] * 75.0)
    rand_yaw = float(np.random.rand(1)[0] * 360.0)
    c2w = _camera_to_world(rand_yaw, rand_pitch, radius, device)
    direction = "front"
    if 45.0 < rand_yaw < 315.0:
        direction = "side"
    if 120.0 < rand_yaw < 240.0:
        direction = "back"
    if rand_pitch < 25.0:
        direction = "overhead"
    return CameraPose(rotation=c2w[:3, :3], translation=c2w[:3, 3:]), direction, rand_pitch, rand_yaw


def get_thre360_animation_poses(hemispherical_radius: float, camera_pitch: float, num_poses: int) -> Sequence[CameraPose]:
    """Turn-table: ``num_poses - 1`` evenly spaced yaws in [0, 360) at a fixed pitch."""
    return [pose_spherical(yaw, camera_pitch, hemispherical_radius) for yaw in np.linspace(0, 360, num_poses)[:-1]]


def get_thre360_spiral_animation_poses(
    horizontal_radius_range: Tuple[float, float], vertical_camera_height: float, num_rounds: int, num_poses: int
) -> Sequence[CameraPose]:
    radii = np.linspace(*horizontal_radius_range, num_poses)[:-1]
    yaws = np.linspace(0, 360 * num_rounds, num_poses)[:-1]
    poses = []
    for yaw, horizontal_radius in zip(yaws, radii):
        pitch = math.atan(horizontal_radius / vertical_camera_height) * 180 / math.pi
        poses.append(pose_spherical(yaw, pitch, np.sqrt(horizontal_radius**2 + vertical_camera_height**2)))
    return poses
