"""Every ``file.py:line`` citation of the reference in the headers, kernels, product Python, oracles and docs must point
into an existing file of the reference tree, inside its length.  Runs only where the reference is present (the build
container); skipped on the GPU box."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
SOURCES = [ROOT / "include" / "voxe.h", ROOT / "DESIGN.md", ROOT / "INTEGRATION.md", *sorted((ROOT / "oracle").glob("*.py")),
           *sorted((ROOT / "vox-e_b200" / "csrc").glob("*.c*")), *sorted((ROOT / "vox-e_b200" / "csrc").glob("*.h")),
           *sorted((ROOT / "vox-e_b200" / "voxe_b200").glob("*.py")), *sorted((ROOT / "vox-e_b200" / "thre3d_atom").rglob("*.py"))]
# path ending in .py, then ":" and one or more comma-separated line numbers / ranges
CITATION = re.compile(r"((?:[\w\-]+/)*[\w\-]+\.py):((?:\d+(?:-\d+)?)(?:,\s?\d+(?:-\d+)?)*)")


def _resolve(rel: str):
    """Citations are written relative to the reference root, to thre3d_atom/, or as a bare file name."""
    for base in (REFERENCE, REFERENCE / "thre3d_atom"):
        if (base / rel).is_file():
            return [base / rel]
    return [p for p in REFERENCE.rglob(Path(rel).name) if str(p).endswith(rel)]


@pytest.mark.skipif(not REFERENCE.is_dir(), reason="reference tree not present")
def test_reference_citations_resolve():
    lengths, checked, bad = {}, 0, []
    for src in SOURCES:
        for rel, spans in CITATION.findall(src.read_text()):
            if rel.startswith(("tests/", "tools/", "oracle/", "voxe_b200/", "vox-e_b200/")) or rel in ("bench.py", "__graft_entry__.py"):
                continue  # our own files
            targets = _resolve(rel)
            if not targets:
                if (ROOT / rel).exists() or list((ROOT / "vox-e_b200").rglob(Path(rel).name)) and "/" not in rel and not list(REFERENCE.rglob(Path(rel).name)):
                    continue  # a file of this repository mentioned with a line number
                bad.append(f"{src.name}: {rel} not found in the reference")
                continue
            last = max(int(n) for n in re.findall(r"\d+", spans))
            ok = False
            for t in targets:  # same-named trainers exist in several modules: one of them must be long enough
                if t not in lengths:
                    lengths[t] = len(t.read_text(errors="replace").splitlines())
                ok |= last <= lengths[t]
            checked += 1
            if not ok:
                bad.append(f"{src.name}: {rel}:{spans} runs past the end of the file")
    assert checked > 100, checked
    assert not bad, "\n".join(bad[:40])
