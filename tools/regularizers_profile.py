"""ncu driver for the row f2 / f3 kernels: three eager calls of each entry point on the benchmark grid (160^3 SH-0),
nothing else.  Numbers printed under a profiler are not bench values."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "vox-e_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from thre3d_atom.utils.imaging_utils import CameraIntrinsics  # noqa: E402
from voxe_b200 import regularizers as reg  # noqa: E402
from voxe_b200 import sampling  # noqa: E402

bench.select_workload("cfg2")
dev = torch.device("cuda:0")
dens, feat = bench.make_grid_tensors(dev)
pre = (dens + 0.1 * torch.randn_like(dens)).contiguous()
dens, feat = torch.nn.Parameter(dens), torch.nn.Parameter(feat)
dens.grad, feat.grad = torch.zeros_like(dens), torch.zeros_like(feat)
poses = torch.eye(3, 4, device=dev).repeat(8, 1, 1).contiguous()
pixels = torch.rand(8 * 800 * 800, 3, device=dev)
for _ in range(3):
    reg.accumulate_density_loss_gradient(dens, pre, 200.0)
    reg.accumulate_tv_gradient(dens, 1.0, relu=True)
    reg.accumulate_tv_gradient(feat, 1.0)
    sampling.sample_rays_from_cameras(CameraIntrinsics(800, 800, 1111.1), poses, pixels, 4096)
torch.cuda.synchronize()
print("done")
