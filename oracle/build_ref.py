#!/usr/bin/env python
"""oracle/build_ref.py -- stage the UNMODIFIED reference's hot-path modules under ``oracle/_ref/`` (test infrastructure).

The reference is pure Python, so "building" it means making the handful of modules on the render path importable where
``/root/reference`` does not exist (the GPU box): this script copies them, byte for byte, from the reference tree into
``oracle/_ref/thre3d_atom/`` and writes two empty stub packages for imports the path never executes
(``matplotlib.pyplot`` at utils/imaging_utils.py:4, used only by ``postprocess_depth_map``; ``easydict`` at
utils/misc.py:6, used only by ``log_config_to_disk``).  ``oracle/_ref/`` is git-ignored (never part of the history) but
travels with the gpurun snapshot, like the built ``.so`` files.

Used by ``bench.py --impl reference`` (the CPU arm then times the reference itself: ``cpu_baseline.kind == "reference"``)
and by bench.py's ``gpu_baseline`` (the reference's own stock-ATen path on the GPU).  Nothing under ``vox-e_b200/`` may
import it; ``__graft_entry__.build()`` runs this when the reference tree is present.

    python oracle/build_ref.py [--reference /root/reference]
"""
import argparse
import hashlib
import json
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"

# the modules `render_sh_voxel_grid` / `VolumetricModel` import, transitively
MODULES = [
    "thre3d_atom/__init__.py",
    "thre3d_atom/utils/__init__.py",
    "thre3d_atom/utils/constants.py",
    "thre3d_atom/utils/imaging_utils.py",
    "thre3d_atom/utils/misc.py",
    "thre3d_atom/rendering/__init__.py",
    "thre3d_atom/rendering/volumetric/__init__.py",
    "thre3d_atom/rendering/volumetric/accumulate.py",
    "thre3d_atom/rendering/volumetric/process.py",
    "thre3d_atom/rendering/volumetric/render_interface.py",
    "thre3d_atom/rendering/volumetric/sample.py",
    "thre3d_atom/rendering/volumetric/utils/__init__.py",
    "thre3d_atom/rendering/volumetric/utils/misc.py",
    "thre3d_atom/rendering/volumetric/utils/spherical_harmonics.py",
    "thre3d_atom/thre3d_reprs/__init__.py",
    "thre3d_atom/thre3d_reprs/constants.py",
    "thre3d_atom/thre3d_reprs/renderers.py",
    "thre3d_atom/thre3d_reprs/voxels.py",
    "thre3d_atom/modules/__init__.py",
    "thre3d_atom/modules/volumetric_model.py",
]
STUBS = {
    "matplotlib/__init__.py": "# stub: the render path never calls into matplotlib (oracle/build_ref.py)\n",
    "matplotlib/pyplot.py": "# stub (oracle/build_ref.py)\n",
    "easydict/__init__.py": "# stub (oracle/build_ref.py)\nclass EasyDict(dict):\n    pass\n",
}


def build(reference: Path) -> dict:
    if OUT.exists():
        shutil.rmtree(OUT)
    manifest = {"reference": str(reference), "files": {}}
    for rel in MODULES:
        src = reference / rel
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        data = src.read_bytes()
        dst.write_bytes(data)
        manifest["files"][rel] = hashlib.sha256(data).hexdigest()[:16]
    for rel, text in STUBS.items():
        dst = OUT / "_stubs" / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_text(text)
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    return manifest


def ref_paths():
    """sys.path entries that make the staged reference importable (the stubs only when the real packages are absent)."""
    import importlib.util

    paths = [str(OUT)]
    if importlib.util.find_spec("matplotlib") is None or importlib.util.find_spec("easydict") is None:
        paths.append(str(OUT / "_stubs"))
    return paths


def available() -> bool:
    return (OUT / "MANIFEST.json").exists()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    m = build(Path(a.reference))
    print(f"[build_ref] staged {len(m['files'])} reference modules under {OUT}")
