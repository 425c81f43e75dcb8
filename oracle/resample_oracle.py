"""oracle/resample_oracle.py -- CPU restatement of the grid rescale of progressive training (test infrastructure: only
tests/ may import it).

``scale_voxel_grid_with_required_output_size`` (thre3d_atom/thre3d_reprs/voxels.py:409-447) concatenates features and
densities, permutes to [1, C, X, Y, Z] and calls ``torch.nn.functional.interpolate(size=output_size, mode="trilinear",
align_corners=False, recompute_scale_factor=False)`` (:421-428).  PyTorch is a third-party dependency of the reference
(pinned ``torch==1.13.0``, requirements.txt:2) whose source is not under /root/reference; the published algorithm of its
``upsample_trilinear3d`` is restated here with explicit index arithmetic (no ``interpolate`` call):

    per axis   scale = in / out;  s = max(scale * (o + 0.5) - 0.5, 0);  i0 = floor(s);  i1 = min(i0 + 1, in - 1);  l1 = s - i0
    value      sum over the 8 (i, j, k) corner combinations of  w_x * w_y * w_z * grid[i, j, k, :]

Pinned against the executed reference: tests/golden/resample.npz (tests/golden/make_golden_resample.py)."""
from typing import Tuple

import numpy as np


def _axis(n_in: int, n_out: int, dtype):
    scale = dtype(n_in) / dtype(n_out)
    s = np.maximum(scale * (np.arange(n_out, dtype=dtype) + dtype(0.5)) - dtype(0.5), dtype(0.0))
    i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = (s - i0.astype(dtype)).astype(dtype)
    return i0, i1, dtype(1.0) - l1, l1


def resample_grid_oracle(grid: np.ndarray, output_size: Tuple[int, int, int], dtype=np.float64) -> np.ndarray:
    """grid [X, Y, Z, C] -> [X2, Y2, Z2, C].  dtype float64: truth; float32: ATen's own precision."""
    g = grid.astype(dtype)
    (x0, x1, wx0, wx1), (y0, y1, wy0, wy1), (z0, z1, wz0, wz1) = (_axis(g.shape[a], output_size[a], dtype) for a in range(3))
    out = np.zeros((*output_size, g.shape[3]), dtype=dtype)
    for xi, wx in ((x0, wx0), (x1, wx1)):
        for yi, wy in ((y0, wy0), (y1, wy1)):
            for zi, wz in ((z0, wz0), (z1, wz1)):
                w = wx[:, None, None] * wy[None, :, None] * wz[None, None, :]
                out += w[..., None] * g[xi[:, None, None], yi[None, :, None], zi[None, None, :], :]
    return out


def rescaled_voxel_size(voxel_size, dims, output_size):
    """voxels.py:434-438: the grid keeps its world extent."""
    return tuple(voxel_size[a] * dims[a] / output_size[a] for a in range(3))
