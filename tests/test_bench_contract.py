"""The reference arm of bench.py (`--impl reference`) runs on the host CPU, so its side of the driver's contract can be
checked without a GPU: one JSON line on stdout with the base keys, `impl: "reference"`, the `cpu_baseline` description of
the run and an `e2e` object repeating the value with zero copy bytes.  (The product arm refuses to run without CUDA.)"""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=280):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=str(ROOT))


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-600:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if (ROOT / "oracle" / "_ref" / "thre3d_atom").is_dir():
        assert cb["kind"] == "reference", "the staged unmodified reference is what the arm must run when it is present"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        return  # on a GPU box the arm runs; the -m gpu tests and the driver cover it
    r = _run("--steps", "1", "--warmup", "0", timeout=120)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")], "no bench line may be printed from a CPU fallback"
