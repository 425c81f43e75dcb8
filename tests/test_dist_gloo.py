"""Multi-process host logic on CPU: world_size-2 gloo run of the ray sharding + single voxel-gradient all-reduce.

The render itself is stood in for by the oracle (this package has no CPU render path, by design); what is under test is
``voxe_b200.dist``: that sharded partial gradients, summed by ONE collective, equal the single-process gradient, for both
sharding schemes, and that the mean-loss rescaling is right with uneven shards.
"""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _case():
    from oracle.voxe_oracle import OracleConfig, OracleGrid, cast_rays_np, pose_spherical_np

    g = torch.Generator().manual_seed(3)
    dims = (10, 10, 10)
    dens = torch.rand((*dims, 1), generator=g) * 2 - 1
    feat = torch.rand((*dims, 3), generator=g) * 2 - 1
    rot, trans = pose_spherical_np(30.0, 55.0, 4.0311)
    rays_o, rays_d = cast_rays_np(7, 9, 9.0, rot, trans)  # 63 rays: odd on purpose (uneven shards)
    target = torch.rand(rays_o.shape[0], 3, generator=g)
    grid = OracleGrid((0.3, 0.3, 0.3), density_scale=6.0, preact="identity", postact="softplus")
    cfg = OracleConfig(num_samples=24, near=1.8, far=6.6, white_bkgd=True)
    return dens, feat, rays_o, rays_d, target, grid, cfg


def _partial_grads(dens, feat, rays_o, rays_d, target, grid, cfg, scale):
    """Gradient of scale * mean((colour - target)^2) over the given rays, through the oracle."""
    from oracle.voxe_oracle import render_oracle

    d = dens.clone().double().requires_grad_(True)
    f = feat.clone().double().requires_grad_(True)
    if rays_o.shape[0] == 0:
        return torch.zeros_like(d), torch.zeros_like(f)
    out = render_oracle(d, f, grid, rays_o, rays_d, cfg)
    loss = ((out["colour"] - target.double()) ** 2).mean() * scale
    loss.backward()
    return d.grad, f.grad


def _worker(rank, world, port, scheme, tmp):
    for p in (ROOT, ROOT / "vox-e_b200"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from voxe_b200 import dist as vd

        dens, feat, rays_o, rays_d, target, grid, cfg = _case()
        n = rays_o.shape[0]
        if scheme == "contiguous":
            lo, hi = vd.shard_bounds(n, rank, world)
            o, d = vd.shard_rays(rays_o, rays_d, rank, world)
            t = target[lo:hi]
        else:
            o, d = vd.shard_rays(rays_o, rays_d, rank, world, batch=8)
            idx = torch.cat([torch.arange(s, min(s + 8, n)) for s in list(range(0, n, 8))[rank::world]])
            t = target[idx]
        scale = vd.global_mean_scale(o.shape[0])
        gd, gf = _partial_grads(dens, feat, o, d, t, grid, cfg, scale)
        pd = torch.nn.Parameter(dens.clone())  # fp32 parameters and gradients, like the CUDA path produces
        pf = torch.nn.Parameter(feat.clone())
        pd.grad, pf.grad = gd.float(), gf.float()
        reducer = vd.VoxelGradAllReducer([pd, pf])
        reducer()
        assert reducer.num_collectives == 1
        torch.save({"d": pd.grad, "f": pf.grad, "n_local": o.shape[0]}, os.path.join(tmp, f"{scheme}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scheme", ["contiguous", "batches"])
def test_sharded_gradients_sum_to_the_single_process_gradient(scheme, tmp_path):
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, scheme, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    dens, feat, rays_o, rays_d, target, grid, cfg = _case()
    want_d, want_f = _partial_grads(dens, feat, rays_o, rays_d, target, grid, cfg, 1.0)
    outs = [torch.load(tmp_path / f"{scheme}_{r}.pt") for r in range(world)]
    assert sum(o["n_local"] for o in outs) == rays_o.shape[0]
    for o in outs:  # every rank ends with the same, full gradient
        assert torch.allclose(o["d"].double(), want_d, rtol=1e-4, atol=1e-9)
        assert torch.allclose(o["f"].double(), want_f, rtol=1e-4, atol=1e-9)


def test_shard_helpers():
    sys.path.insert(0, str(ROOT / "vox-e_b200"))
    from voxe_b200 import dist as vd

    assert [vd.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [vd.shard_bounds(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert vd.shard_views(100, 3, 8) == list(range(3, 100, 8)) and len(vd.shard_views(100, 3, 8)) == 13
    with pytest.raises(ValueError):
        vd.shard_bounds(10, 4, 4)
    o = torch.arange(30.0).reshape(10, 3)
    a, _ = vd.shard_rays(o, o, 1, 2, batch=4)  # batches [0:4] [4:8] [8:10] -> rank 1 gets [4:8]
    assert torch.equal(a, o[4:8])
    assert vd.world_info() == (0, 1)
    assert vd.global_mean_scale(5) == 1.0


class _StubPeerVolume:
    """Stands in for PeerGradVolume on the CPU: same surface, the exchange done by gloo."""

    def __init__(self, buffer):
        self.buffer, self.calls = buffer, []

    def allreduce(self):
        self.calls.append("dense")
        dist.all_reduce(self.buffer)

    def allreduce_sparse(self, tag):
        self.calls.append(("sparse", tag))
        dist.all_reduce(self.buffer)


class _StubGrid:
    def __init__(self, acc, params):
        self.render_gradient_accumulator = acc
        self.densities, self.features = params


def _deferred_worker(rank, world, port, tmp):
    for p in (ROOT, ROOT / "vox-e_b200"):
        sys.path.insert(0, str(p))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from voxe_b200 import dist as vd
        from voxe_b200.render_function import PackedGradAccumulator

        params = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(4))]
        record = {}
        for mode in ("flat", "peer_dense", "peer_sparse"):
            acc = PackedGradAccumulator()
            acc.buffer = torch.full((16,), float(rank + 1))
            if mode != "flat":
                acc.peer_volume = _StubPeerVolume(acc.buffer)
            if mode == "peer_sparse":
                acc.sparse_sink = True
                acc.touch_tag[0] = 7
            reducer = vd.VoxelGradAllReducer(params, grids=[_StubGrid(acc, params)])
            assert reducer.reduce_deferred() == 1 and reducer.num_collectives == 1
            record[mode] = (acc.buffer.clone(), None if mode == "flat" else list(acc.peer_volume.calls), acc.dirty)
        # a sink that was re-allocated (no longer the peer's buffer) must not be exchanged through the peer volume
        acc = PackedGradAccumulator()
        acc.buffer = torch.full((16,), float(rank + 1))
        acc.peer_volume = _StubPeerVolume(torch.zeros(16))
        vd.VoxelGradAllReducer(params, grids=[_StubGrid(acc, params)]).reduce_deferred()
        record["stale_peer"] = (acc.buffer.clone(), list(acc.peer_volume.calls), acc.dirty)
        torch.save(record, os.path.join(tmp, f"deferred_{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_reduce_deferred_takes_the_peer_volume_when_the_sink_lives_in_it(tmp_path):
    """Host logic of the deferred-gradient exchange at world size 2: the packed sink is reduced through the library's peer
    kernel when it IS the peer-mapped volume (brick-wise when the sink keeps a trail, with the step's tag), through
    torch.distributed otherwise; the sink is marked dirty so that the optimiser-side hand-over runs on every rank."""
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_deferred_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        rec = torch.load(tmp_path / f"deferred_{r}.pt")
        for mode, calls in (("flat", None), ("peer_dense", ["dense"]), ("peer_sparse", [("sparse", 7)]), ("stale_peer", [])):
            buf, got_calls, dirty = rec[mode]
            assert torch.equal(buf, torch.full((16,), 3.0)), (r, mode)  # 1 + 2 on both ranks
            assert got_calls == calls and dirty, (r, mode, got_calls)
