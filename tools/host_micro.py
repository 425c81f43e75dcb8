"""Host-side cost of individual torch calls used by the e2e path (run on a GPU box)."""
import time
import torch

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def t(name, fn, n=200, sync_each=False):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        if sync_each:
            torch.cuda.synchronize()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:50s} host {1e6 * (t1 - t0) / n:8.1f} us/call   drained {1e6 * (t2 - t0) / n:8.1f} us/call", flush=True)


buf = torch.empty(4096, 256, device=dev)
t("torch.rand(4096,256)", lambda: torch.rand(4096, 256, dtype=torch.float32, device=dev))
t("torch.rand(4096,256) sync each", lambda: torch.rand(4096, 256, dtype=torch.float32, device=dev), sync_each=True)
t("buf.uniform_()", lambda: buf.uniform_())
t("torch.empty(4096,256)", lambda: torch.empty(4096, 256, device=dev))
t("torch.empty(4096,256).uniform_()", lambda: torch.empty(4096, 256, device=dev).uniform_())
t("torch.rand(512,256)", lambda: torch.rand(512, 256, dtype=torch.float32, device=dev))
t("torch.randn(4096,3)", lambda: torch.randn(4096, 3, device=dev))
g = torch.Generator(device=dev)
t("torch.rand(4096,256, generator=g)", lambda: torch.rand(4096, 256, dtype=torch.float32, device=dev, generator=g))
h = torch.empty(4096, 3).pin_memory()
t("pinned.to(dev, non_blocking)", lambda: h.to(dev, non_blocking=True))
x = torch.empty(4096, 3, device=dev)
t("mul+sum", lambda: (x * x).sum())
