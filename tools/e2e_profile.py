"""cProfile of bench.py's e2e leg (host-side overhead of the public API path); run on a GPU box."""
import cProfile
import io
import pstats
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402

bench.select_workload("cfg2")
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
deferred = len(sys.argv) > 1 and sys.argv[1] == "deferred"
bench.e2e_leg(dev, 0, 1, 2, 2, None, deferred=deferred)  # warm
pr = cProfile.Profile()
pr.enable()
r = bench.e2e_leg(dev, 0, 1, 6, 1, None, deferred=deferred)
pr.disable()
print(r)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue())
