"""Golden vectors for the per-step grid regularisers, produced by EXECUTING the reference's own function bodies.

    python tests/golden/make_golden_regularizers.py       # rewrites tests/golden/regularizers.npz

``thre3d_atom/modules/sds_trainer.py`` cannot be imported here (it pulls in diffusers / wandb / the dataset stack), so
the four functions are cut out of the unmodified file with ``ast`` and executed verbatim in a namespace that holds what
the file itself imports for them (``torch``, ``mse_loss``, ``l1_loss``, sds_trainer.py:10-12).  Run in the build
container only: the reference tree does not exist on the GPU box.

Each case stores the inputs and, from the reference in fp32 on CPU: the loss, upstream * dloss/dinput (autograd) and,
for the correlation loss, its second return value.
"""
import ast
import json
from pathlib import Path

import numpy as np
import torch
from torch.nn.functional import l1_loss, mse_loss

REFERENCE_FILE = "/root/reference/thre3d_atom/modules/sds_trainer.py"
WANTED = ("density_correlation_loss_fn", "_density_correlation_loss", "_tv_loss_on_grid")
OUT = Path(__file__).resolve().parent / "regularizers.npz"


def reference_functions():
    source = Path(REFERENCE_FILE).read_text()
    tree = ast.parse(source)
    namespace = {"torch": torch, "Tensor": torch.Tensor, "mse_loss": mse_loss, "l1_loss": l1_loss}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REFERENCE_FILE, "exec"), namespace)
    return {name: namespace[name] for name in WANTED}


def grids(rng, dims, channels, kind):
    x = torch.from_numpy(rng.standard_normal((*dims, channels)).astype(np.float32))
    if kind == "smooth":  # low-frequency content plus noise, some exactly equal neighbours (sign(0) = 0 in the TV gradient)
        X, Y, Z = dims
        gx, gy, gz = np.meshgrid(np.linspace(-1, 1, X), np.linspace(-1, 1, Y), np.linspace(-1, 1, Z), indexing="ij")
        base = np.cos(2.5 * gx) * np.sin(1.7 * gy + 0.3) + gz
        x = 0.2 * x + torch.from_numpy(base.astype(np.float32))[..., None]
        x[: X // 2, : Y // 2] = torch.round(x[: X // 2, : Y // 2] * 2) / 2  # plateaus: ties between neighbours
    return x.contiguous()


def main():
    ref = reference_functions()
    rng = np.random.default_rng(20261017)
    out, meta = {}, {"tv": [], "pair": []}

    tv_cases = [("dens_relu", (9, 7, 11), 1, True, "smooth", 1.0), ("dens_plain", (8, 8, 8), 1, False, "noise", 1.0),
                ("feat3", (6, 10, 5), 3, False, "smooth", 0.37), ("feat27", (4, 5, 6), 27, False, "noise", 2.5),
                ("thin", (2, 2, 33), 4, True, "noise", 1.0), ("attn", (12, 12, 12), 1, False, "smooth", 0.01)]
    for name, dims, channels, relu, kind, upstream in tv_cases:
        x = grids(rng, dims, channels, kind).requires_grad_(True)
        loss = ref["_tv_loss_on_grid"](torch.nn.ReLU()(x) if relu else x)  # sds_trainer.py:319-321
        (loss * upstream).backward()
        out[f"tv_{name}_grid"], out[f"tv_{name}_loss"], out[f"tv_{name}_grad"] = x.detach().numpy(), loss.detach().numpy(), x.grad.numpy()
        meta["tv"].append({"name": name, "relu": relu, "upstream": upstream})

    pair_cases = [("corr_near", (10, 9, 8), "correlation", 0.05, 200.0), ("corr_far", (7, 7, 7), "correlation", 1.5, 1.0),
                  ("l2", (6, 5, 9), "l2", 0.7, 3.0), ("l1", (6, 5, 9), "l1", 0.7, 0.5)]
    for name, dims, mode, noise, upstream in pair_cases:
        b = grids(rng, dims, 1, "smooth") * 20.0
        a = (b + noise * 20.0 * torch.from_numpy(rng.standard_normal(b.shape).astype(np.float32))).requires_grad_(True)
        if mode != "correlation":
            a.data[0, 0, :3] = b[0, 0, :3]  # exact ties: sign(0) = 0 in the L1 gradient
        loss, grid = ref["density_correlation_loss_fn"](sds_density=a, regular_density=b, l2_mode=mode == "l2", l1_mode=mode == "l1")
        (loss * upstream).backward()
        out[f"pair_{name}_a"], out[f"pair_{name}_b"] = a.detach().numpy(), b.numpy()
        out[f"pair_{name}_loss"], out[f"pair_{name}_grad"] = loss.detach().numpy(), a.grad.numpy()
        if grid is not None:
            out[f"pair_{name}_corr"] = grid.numpy()
        meta["pair"].append({"name": name, "mode": mode, "upstream": upstream})

    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes, {len(tv_cases)} TV cases, {len(pair_cases)} pair cases)")


if __name__ == "__main__":
    main()
