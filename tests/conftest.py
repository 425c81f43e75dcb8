"""pytest configuration: registers the ``gpu`` marker and puts the product tree (``vox-e_b200/``) on sys.path.

``vox-e_b200/`` is a path root, not an importable name (hyphen): it carries ``voxe_b200`` (native binding),
``thre3d_atom`` (the reference-facing interface for the hot path) and ``csrc`` (CUDA + C-ABI sources).
"""
import os
import sys
from pathlib import Path

import pytest

REPO_ROOT = Path(__file__).resolve().parent.parent
for p in (REPO_ROOT, REPO_ROOT / "vox-e_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    """Same autouse seeding as the reference's conftest (thre3d_atom/conftest.py:18-21)."""
    import numpy as np
    import torch

    torch.manual_seed(42)
    np.random.seed(42)
