"""All-reduce time of the packed gradient volume (68 MB fp32) over the box's GPUs; run under torchrun."""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
for n in (17006112, 17006112 * 4):
    x = torch.randn(n, device="cuda")
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        dist.all_reduce(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = n * 4 / 1e9
        print(f"{os.environ.get('TAG', 'default'):40s} world {world}  {gb * 1e3:7.1f} MB  {float(t):7.3f} ms  algbw {gb / float(t) * 1e3:7.1f} GB/s  busbw {gb / float(t) * 1e3 * 2 * (world - 1) / world:7.1f} GB/s", flush=True)
dist.destroy_process_group()
