"""Golden vectors for the stand-alone point query ``VoxelGrid.forward`` / ``forward_attn`` (thre3d_reprs/voxels.py:287-345,
347-406 upstream), produced by EXECUTING the unmodified reference on CPU (build container only; stubs as in make_golden.py):

    python tests/golden/make_golden_points.py      # rewrites tests/golden/points.npz

Points are drawn over 1.6x the grid's extent (so some fall in the half-voxel border, some outside it, some far away), plus
the eight box corners, the centre and points exactly on faces.  Stored per case: grid tensors, points, the reference's
output rows, and d(sum(out * G))/d(densities, features[, attn]) from its autograd for a stored random G.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT_DIR, _import_reference  # noqa: E402

ACT = {"identity": lambda: torch.nn.Identity(), "abs": lambda: torch.abs, "relu": lambda: torch.nn.ReLU(), "softplus": lambda: torch.nn.Softplus()}

CASES = [
    dict(name="default_abs_identity", dims=(5, 6, 7), n_feat=3, pre="abs", post="identity", scale=1.0, voxel=(0.3, 0.25, 0.2), loc=(0.0, 0.0, 0.0)),
    dict(name="relu_field", dims=(8, 8, 8), n_feat=3, pre="identity", post="relu", scale=33.333, voxel=(0.25, 0.25, 0.25), loc=(0.0, 0.0, 0.0)),
    dict(name="softplus_offcentre", dims=(4, 9, 3), n_feat=3, pre="identity", post="softplus", scale=5.0, voxel=(0.2, 0.1, 0.4), loc=(0.3, -0.2, 0.1)),
    dict(name="sh1_abs_relu", dims=(6, 5, 4), n_feat=12, pre="abs", post="relu", scale=2.0, voxel=(0.3, 0.3, 0.3), loc=(0.0, 0.1, 0.0)),
    dict(name="sh2", dims=(5, 5, 6), n_feat=27, pre="identity", post="relu", scale=33.333, voxel=(0.3, 0.3, 0.25), loc=(0.0, 0.0, 0.0)),
    dict(name="sh3_thin", dims=(1, 4, 5), n_feat=48, pre="identity", post="identity", scale=1.0, voxel=(0.5, 0.3, 0.2), loc=(0.0, 0.0, 0.0)),
    dict(name="attn", dims=(6, 6, 6), n_feat=3, pre="identity", post="relu", scale=33.333, voxel=(0.3, 0.3, 0.3), loc=(0.0, 0.0, 0.0), attn=True),
    dict(name="attn_orig_densities", dims=(6, 6, 6), n_feat=3, pre="identity", post="relu", scale=33.333, voxel=(0.3, 0.3, 0.3), loc=(0.0, 0.0, 0.0),
         attn=True, orig=True),
]
N_RANDOM = 300


def _points(case, g):
    ext = torch.tensor([n * v for n, v in zip(case["dims"], case["voxel"])])
    loc = torch.tensor(case["loc"])
    pts = loc + (torch.rand(N_RANDOM, 3, generator=g) - 0.5) * 1.6 * ext
    lo, hi = loc - ext / 2, loc + ext / 2
    corners = torch.stack([torch.stack([(lo, hi)[(k >> a) & 1][a] for a in range(3)]) for k in range(8)])
    faces = loc.repeat(4, 1)
    faces[0, 0], faces[1, 1], faces[2, 2] = lo[0], hi[1], lo[2]   # exactly on three faces; faces[3] = the centre
    far = loc + torch.tensor([[5.0, 0.0, 0.0], [0.0, -7.0, 0.0], [3.0, 3.0, 3.0]]) * ext
    return torch.cat([pts, corners, faces, far]).float().contiguous()


def main():
    ref = _import_reference()
    arrays, meta = {}, {}
    for k, case in enumerate(CASES):
        g = torch.Generator().manual_seed(500 + k)
        dims = case["dims"]
        dens = (torch.randn((*dims, 1), generator=g) * 0.5).requires_grad_(True)
        feat = torch.randn((*dims, case["n_feat"]), generator=g).requires_grad_(True)
        attn = torch.randn((*dims, 1), generator=g).requires_grad_(True) if case.get("attn") else None
        grid = ref["VoxelGrid"](dens, feat, ref["VoxelSize"](*case["voxel"]), ref["VoxelGridLocation"](*case["loc"]),
                                density_preactivation=ACT[case["pre"]](), density_postactivation=ACT[case["post"]](),
                                expected_density_scale=case["scale"], tunable=False, **({"attn": attn} if attn is not None else {}))
        orig = None
        if case.get("orig"):
            orig = (torch.randn((*dims, 1), generator=g) * 0.5).requires_grad_(True)
            grid.orig_densities = orig
        pts = _points(case, g)
        out = grid.forward_attn(pts, orig_densities=bool(case.get("orig"))) if case.get("attn") else grid(pts)
        G = torch.randn(out.shape, generator=g)
        (out * G).sum().backward()
        n = case["name"]
        arrays[f"{n}/densities"], arrays[f"{n}/features"] = dens.detach().numpy(), feat.detach().numpy()
        arrays[f"{n}/points"], arrays[f"{n}/out"], arrays[f"{n}/g_out"] = pts.numpy(), out.detach().numpy(), G.numpy()
        if case.get("attn"):
            arrays[f"{n}/attn"], arrays[f"{n}/d_attn"] = attn.detach().numpy(), attn.grad.numpy()
            src = orig if orig is not None else dens
            arrays[f"{n}/d_densities"] = src.grad.numpy()
            if orig is not None:
                arrays[f"{n}/orig_densities"] = orig.detach().numpy()
        else:
            arrays[f"{n}/d_densities"], arrays[f"{n}/d_features"] = dens.grad.numpy(), feat.grad.numpy()
        meta[n] = {key: case[key] for key in ("dims", "n_feat", "pre", "post", "scale", "voxel", "loc")} | {
            "attn": bool(case.get("attn")), "orig": bool(case.get("orig")), "n_points": int(pts.shape[0])}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT_DIR / "points.npz", **arrays)
    print(f"wrote {OUT_DIR / 'points.npz'} ({len(CASES)} cases)")


if __name__ == "__main__":
    main()
