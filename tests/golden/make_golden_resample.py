"""Golden vectors for ``scale_voxel_grid_with_required_output_size`` (thre3d_reprs/voxels.py:409-447 upstream), produced by
EXECUTING the unmodified reference on CPU (build container only; same two stub modules as make_golden.py):

    python tests/golden/make_golden_resample.py      # rewrites tests/golden/resample.npz

Cases: up-scaling by a non-integer factor on an anisotropic grid, down-scaling, identity size, an extent-1 axis, and the
SH-2 channel count.  Stored per case: input densities / features, the output size, the reference's output tensors and
its new voxel size.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT_DIR, _import_reference  # noqa: E402

CASES = [
    dict(name="up_aniso", dims=(5, 6, 7), out=(8, 9, 11), n_feat=3),
    dict(name="down", dims=(9, 7, 8), out=(6, 5, 4), n_feat=3),
    dict(name="same", dims=(4, 4, 4), out=(4, 4, 4), n_feat=3),
    dict(name="thin_axis", dims=(1, 6, 5), out=(3, 4, 9), n_feat=3),
    dict(name="sh2_double", dims=(6, 5, 4), out=(12, 10, 8), n_feat=27),
]


def main():
    ref = _import_reference()
    from thre3d_atom.thre3d_reprs.voxels import scale_voxel_grid_with_required_output_size

    arrays, meta = {}, {}
    for k, case in enumerate(CASES):
        g = torch.Generator().manual_seed(100 + k)
        dens = torch.randn((*case["dims"], 1), generator=g)
        feat = torch.randn((*case["dims"], case["n_feat"]), generator=g)
        grid = ref["VoxelGrid"](dens.clone(), feat.clone(), ref["VoxelSize"](0.3, 0.25, 0.2), tunable=False)
        new = scale_voxel_grid_with_required_output_size(grid, case["out"])
        n = case["name"]
        arrays[f"{n}/densities"], arrays[f"{n}/features"] = dens.numpy(), feat.numpy()
        arrays[f"{n}/out_densities"] = new.densities.detach().contiguous().numpy()
        arrays[f"{n}/out_features"] = new.features.detach().contiguous().numpy()
        meta[n] = {"dims": case["dims"], "out": case["out"], "n_feat": case["n_feat"], "voxel_size": [0.3, 0.25, 0.2],
                   "new_voxel_size": [float(v) for v in new.voxel_size]}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT_DIR / "resample.npz", **arrays)
    print(f"wrote {OUT_DIR / 'resample.npz'} ({len(CASES)} cases)")


if __name__ == "__main__":
    main()
