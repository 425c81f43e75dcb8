"""Time the regulariser kernels alone (the `regularizers` object of the bench line) without running the whole bench."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "vox-e_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

bench.select_workload("cfg2")
print(json.dumps(bench.regularizers_leg(torch.device("cuda:0"), bench.measured_hbm_peak()[0])))
