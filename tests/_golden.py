"""Helpers shared by the parity tests: load a golden case (tests/golden/*.npz, produced by the executed
reference) and run the oracle on it."""
import json
from pathlib import Path

import numpy as np
import torch

from oracle.voxe_oracle import OracleConfig, OracleGrid, render_oracle_with_grads

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
RENDER_CASES = sorted(p.stem[len("render_"):] for p in GOLDEN_DIR.glob("render_*.npz"))


def load_case(name: str):
    z = np.load(GOLDEN_DIR / f"render_{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def load_npz(name: str):
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else {}
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}


def oracle_grid_cfg(meta):
    grid = OracleGrid(
        voxel_size=tuple(meta["voxel_size"]), location=tuple(meta["location"]), density_scale=meta["density_scale"],
        preact=meta["preact"], postact=meta["postact"],
    )
    cfg = OracleConfig(
        num_samples=meta["num_samples"], near=meta["near"], far=meta["far"], perturb=meta["perturb"],
        optimized_sampling=meta["optimized_sampling"], linear_disparity_sampling=meta["linear_disparity_sampling"],
        white_bkgd=meta["white_bkgd"], render_diffuse=meta["render_diffuse"],
    )
    return grid, cfg


def run_oracle_on_case(meta, a, dtype=torch.float64):
    grid, cfg = oracle_grid_cfg(meta)
    return render_oracle_with_grads(
        a["densities"], a["features"], grid, a["rays_o"], a["rays_d"], cfg,
        g_colour=a["g_colour"], g_depth=a.get("g_depth"), g_acc=a.get("g_acc"), g_disp=a.get("g_disp"),
        jitter=a.get("jitter"), dtype=dtype,
    )


def grad_errors(got: torch.Tensor, want: torch.Tensor):
    """(relative L2, max-abs / ||want||_inf) -- the two gradient figures SURVEY.md 8c sets tolerances on."""
    got, want = got.double(), want.double()
    denom_l2 = want.norm().item() or 1.0
    denom_inf = want.abs().max().item() or 1.0
    return (got - want).norm().item() / denom_l2, (got - want).abs().max().item() / denom_inf
