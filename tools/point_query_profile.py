"""ncu driver for the stand-alone point query: 2^20 points on the 160^3 benchmark grid (SH-0) and on a 96^3 SH-2 grid, forward
and backward, three times each.  Numbers printed under a profiler are not bench values."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "vox-e_b200")]
import torch  # noqa: E402

from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for n, f in ((160, 3), (96, 27)):
    grid = VoxelGrid((torch.randn((n, n, n, 1), generator=g) * 0.01).to(dev), torch.randn((n, n, n, f), generator=g).to(dev),
                     VoxelSize(*(3.0 / n,) * 3), density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                     expected_density_scale=33.333, tunable=True)
    pts = ((torch.rand((1 << 20, 3), generator=g) - 0.5) * 3.3).to(dev)
    gout = torch.randn((1 << 20, f + 1), generator=g).to(dev)
    for _ in range(3):
        (grid(pts) * gout).sum().backward()
torch.cuda.synchronize()
print("done")
