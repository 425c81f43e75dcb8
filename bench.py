#!/usr/bin/env python
"""bench.py -- rays/s forward+backward of the fused ray-marcher on BASELINE.json's headline workload.

Workload (``configs[1]`` of BASELINE.json, SURVEY.md 8d "cfg 2"): 160^3 SH-0 ReLU-field grid, U(-1,1) values (seed 42),
density scale 33.333, world box [-1.5,1.5]^3; 400x400 pinhole camera f=555.5 on the r=4.0311 turn-table
(``get_thre360_animation_poses(4.0311, 60, 9)`` -> 8 poses); S=256 samples on [1.8, 6.6], stratified jitter on, white
background; every frame is rendered as 40 consecutive flat-index batches of <= 4096 rays, each batch forward AND backward
(upstream gradient = a fixed dense dL/dcolour, i.e. loss = <colour, G>).

A *step* is one frame = 160 000 rays = 40 x (forward kernel, backward kernel accumulating into the packed gradient
volume; the stratified jitter is generated inside both kernels, ``--jitter buffer`` draws it with torch.rand instead)
+ one gradient zero-fill + one unpack of the packed gradient into d_densities / d_features (+ ONE all-reduce of the packed
gradient when N > 1: each rank renders its own pose -- weak scaling; the library's own peer-memory kernel
``voxe_allreduce_grads_peer``, ``--collective nccl`` for ncclAllReduce through torch.distributed).

  value     device-resident throughput: the step above replayed as a CUDA graph over C-ABI launches, inputs in HBM;
            three batches are in flight on three streams (--lanes 3; the strictly serialised number is reported as
            ``serialized``); ``value_softplus`` is the same frame on a Softplus field (every in-grid sample scatters)
  parity    the timed leg's own outputs (colours and voxel gradients of the last timed frame) against the oracle run in
            fp32 on the same GPU with the very jitter the kernels drew (voxe_jitter_fill)
  e2e       the same frame through the public API (``VolumetricModel.render_rays`` + ``.backward()`` per 4096-ray batch,
            torch's default autograd engine) with rays and upstream gradients starting in pinned HOST memory and loss +
            colour read back every step; variants (CUDA-graph capture of the same calls, calling-thread engine, deferred
            gradients, one call per frame) beside it
  roofline  the backward kernel (dominant) timed alone with CUDA events: the bytes it moves (counted by the kernel
            itself: saved-vector reloads + the scatters it really issues) / duration vs measured HBM peak; the SURVEY 8d
            contract model beside it as ``model_frac``; L2-side figure from the committed ncu capture
  gpu_baseline  the same batch through stock ATen ops on the same GPU (the oracle, or the reference itself when
            ``oracle/_ref`` was built), CUDA events
  cpu_baseline  the reference (``oracle/_ref``) or its oracle port (PyTorch fp32, all host threads) on a bounded sample

``--impl reference`` times only the CPU leg, one 4096-ray batch forward+backward per step.
``--check`` (any N): ranks render disjoint ray shards, all-reduce with each collective, compare with the unsharded gradient.
``--workload cfg4|cfg5 [--gpus N]``: the other BASELINE.json configurations as stated there (cfg 4: 100 ``get_random_pose``
views dealt round-robin over the ranks, one all-reduce per step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (ROOT, ROOT / "vox-e_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ---------------------------------------------------------------------------------------------------------
# workload definition (cfg 2)
# ---------------------------------------------------------------------------------------------------------
WL = dict(
    name="cfg2: 160^3 SH-0 grid, 400x400 render, 4096-ray batches fwd+bwd, S=256",
    dims=(160, 160, 160), sh_degree=0, world=(3.0, 3.0, 3.0), density_scale=33.333, preact="identity", postact="relu",
    height=400, width=400, focal=555.5, radius=4.0311, pitch=60.0, num_poses=9, S=256, near=1.8, far=6.6,
    batch=4096, perturb=True, white_bkgd=True, seed=42,
)
# Bytes per in-grid sample.  MODEL_* is the SURVEY.md 8d contract (three cache-less corner sweeps: forward gather, backward
# re-gather, scatter payload) kept as ``model_frac``; MOVED_* is what the kernels really move: the forward gathers 8 corners
# and saves one 16-byte vector, the backward reloads that vector and scatters 8 corners only for the samples whose
# gradient is non-zero (the fraction is counted by the kernel itself, VoxeRenderDesc.stats).
MODEL_BYTES_PER_SAMPLE_FWD = 8 * 4 * 4
MODEL_BYTES_PER_SAMPLE_BWD = 2 * 8 * 4 * 4
MODEL_BYTES_PER_RAY = 24 + 24
MOVED_GATHER = 8 * 4 * 4                  # 8 corners x roundup4(F+1) channels x 4 B
MOVED_SAVED_VECTOR = 16
N_SEGMENTS = 16                           # depth segments per ray at S >= 128 (pick_shape in csrc/voxe_capi.cu)


def moved_bytes_per_ray(direction):
    """Per-ray traffic outside the sample loop: rays, outputs / upstream gradients, segment summaries ((n_colour + 3) floats
    per ray and depth segment, written by the forward and read by the backward)."""
    summaries = (3 + 3) * N_SEGMENTS * 4
    return 24 + (12 + 4 + 4 + 4 if direction == "fwd" else 12) + summaries


# The other BASELINE.json configurations (parity cases in tests/test_baseline_configs.py); `--workload cfgN` times their
# device-resident leg for the record -- the bench line the driver reads is always cfg 2.
OTHER_WORKLOADS = {
    "cfg3": dict(name="cfg3: 160^3 SH-2 grid, 512x512 render, one 262144-ray differentiable batch, S=256", sh_degree=2,
                 postact="softplus", height=512, width=512, focal=711.1, batch=512 * 512),
    "cfg4": dict(name="cfg4: 256^3 SH-0 grid, 100 random 800x800 views sharded over the ranks, one view per launch, S=256", dims=(256, 256, 256),
                 postact="softplus", height=800, width=800, focal=1111.1, batch=800 * 800, random_views=100, grid_copies=1),
    "cfg5": dict(name="cfg5: 512^3 SH-2 grid (15 GB), 1024x1024 render, S=512, 65536-ray batches", dims=(512, 512, 512), sh_degree=2,
                 postact="softplus", height=1024, width=1024, focal=1422.2, S=512, batch=65536, grid_copies=1),
}


def select_workload(name):
    global MODEL_BYTES_PER_SAMPLE_FWD, MODEL_BYTES_PER_SAMPLE_BWD, MOVED_GATHER, N_SEGMENTS
    if name != "cfg2":
        WL.update(OTHER_WORKLOADS[name])
    ch = 3 * (WL["sh_degree"] + 1) ** 2 + 1
    MODEL_BYTES_PER_SAMPLE_FWD = 8 * ch * 4
    MODEL_BYTES_PER_SAMPLE_BWD = 2 * 8 * ch * 4
    MOVED_GATHER = 8 * ((ch + 3) // 4) * 16
    N_SEGMENTS = (WL["S"] + 15) // 16


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    return 6650.0, "of fallback (B200_PROFILING.md)"


def make_grid_tensors(device):
    g = torch.Generator().manual_seed(WL["seed"])
    dens = (torch.rand((*WL["dims"], 1), generator=g) * 2 - 1).to(device)
    n_feat = 3 * (WL["sh_degree"] + 1) ** 2
    if dens.numel() * n_feat > 2**29:  # the 15 GB grid: draw on the device
        gg = torch.Generator(device=device).manual_seed(WL["seed"])
        return dens, torch.rand((*WL["dims"], n_feat), device=device, generator=gg) * 2 - 1
    feat = (torch.rand((*WL["dims"], n_feat), generator=g) * 2 - 1).to(device)
    return dens, feat


def make_poses():
    from thre3d_atom.utils.imaging_utils import get_thre360_animation_poses

    return get_thre360_animation_poses(WL["radius"], WL["pitch"], WL["num_poses"])


def frame_rays(pose, device):
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics

    rays = flatten_rays(cast_rays(CameraIntrinsics(WL["height"], WL["width"], WL["focal"]), pose, device=device))
    return rays.origins.contiguous(), rays.directions.contiguous()


def count_inside_samples(rays_o, rays_d):
    """S_in as SURVEY.md 8d defines it: in-AABB samples at the un-jittered depths (strict compare), per frame."""
    t = torch.linspace(0.0, 1.0, WL["S"], device=rays_o.device)
    z = WL["near"] * (1.0 - t) + WL["far"] * t
    half = [w / 2 for w in WL["world"]]
    total = 0
    per_batch = []
    for s in range(0, rays_o.shape[0], WL["batch"]):
        o, d = rays_o[s : s + WL["batch"]], rays_d[s : s + WL["batch"]]
        pts = o[:, None, :] + d[:, None, :] * z[None, :, None]
        inside = torch.ones(pts.shape[:2], dtype=torch.bool, device=pts.device)
        for a in range(3):
            inside &= (pts[..., a] > -half[a]) & (pts[..., a] < half[a])
        n = int(inside.sum().item())
        per_batch.append(n)
        total += n
    return total, per_batch


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and clock-event (throttle) reasons of one GPU through NVML every few ms while the timed
    region runs (falls back to ``nvidia-smi -lms`` when the NVML binding is unavailable)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.004):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.power = [], set(), []
        self.smax, self._stop, self.thread, self.proc, self.lines = None, threading.Event(), None, None, []
        self.nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001 -- no NVML: use the CLI
            self.nvml = None
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except OSError:
                self.proc = None

    def _poll(self):
        n = self.nvml
        masks = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, m in masks.items():
                    if bits & m:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.smax,
                    "power_w_max": max(self.power) if self.power else None, "samples": len(self.samples), "reasons": sorted(self.reasons),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference itself (oracle/_ref, staged by oracle/build_ref.py) or its oracle port, stock ATen ops
# ---------------------------------------------------------------------------------------------------------
def reference_kind():
    from oracle import build_ref

    return "reference" if build_ref.available() else "port"


def use_staged_reference():
    """Make ``thre3d_atom`` resolve to the staged reference instead of this repository's mirror (reference arm only: the
    two packages share their import paths, so one process can hold only one of them)."""
    from oracle import build_ref

    product = str(ROOT / "vox-e_b200")
    sys.path[:] = [q for q in sys.path if q != product]
    for q in reversed(build_ref.ref_paths()):
        sys.path.insert(0, q)
    for name in [m for m in sys.modules if m == "thre3d_atom" or m.startswith("thre3d_atom.")]:
        del sys.modules[name]


def reference_leg(device, steps, warmup, budget_s=None):
    """One step = one 4096-ray batch of pose 0 forward+backward, fp32, through the reference's own render procedure
    (``render_sh_voxel_grid`` of the staged reference; kind "reference") or, when ``oracle/_ref`` was not built, through
    the oracle port (kind "port").  ``device`` cpu: all host threads, perf_counter; cuda: CUDA events."""
    kind = reference_kind()
    on_gpu = device.type == "cuda"
    threads = os.cpu_count() or 1
    if not on_gpu:
        torch.set_num_threads(threads)
    if kind == "reference":
        use_staged_reference()
    dens, feat = make_grid_tensors(device)
    rays_o, rays_d = frame_rays(make_poses()[0], device)
    voxel = tuple(w / d for w, d in zip(WL["world"], WL["dims"]))
    if kind == "reference":
        from thre3d_atom.rendering.volumetric.render_interface import Rays
        from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
        from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
        from thre3d_atom.utils.imaging_utils import CameraBounds

        post = torch.nn.ReLU() if WL["postact"] == "relu" else torch.nn.Softplus()
        grid = VoxelGrid(dens, feat, VoxelSize(*voxel), density_preactivation=torch.nn.Identity(), density_postactivation=post,
                         expected_density_scale=WL["density_scale"], tunable=True)
        cfg = SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]), white_bkgd=True,
                                    perturb_sampled_points=True)

        def one(o, d, gcol):
            grid.densities.grad = None
            grid.features.grad = None
            render_sh_voxel_grid(grid, Rays(o, d), cfg).colour.backward(gcol)
    else:
        from oracle.voxe_oracle import OracleConfig, OracleGrid, render_oracle_with_grads

        ogrid = OracleGrid(voxel, density_scale=WL["density_scale"], preact=WL["preact"], postact=WL["postact"])
        ocfg = OracleConfig(num_samples=WL["S"], near=WL["near"], far=WL["far"], perturb=True, white_bkgd=True)

        def one(o, d, gcol):
            jitter = torch.rand(o.shape[0], WL["S"], device=device)
            render_oracle_with_grads(dens, feat, ogrid, o, d, ocfg, gcol, jitter=jitter, dtype=torch.float32)

    g = torch.Generator().manual_seed(1)
    B = WL["batch"]
    n_batches = rays_o.shape[0] // B
    times = []
    t_begin = time.perf_counter()
    for k in range(warmup + steps):
        b = (17 + k) % n_batches  # start mid-frame so the batches cross the object
        o, d = rays_o[b * B : (b + 1) * B], rays_d[b * B : (b + 1) * B]
        gcol = torch.randn(B, 3, generator=g).to(device)
        if on_gpu:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            e0.record()
            one(o, d, gcol)
            e1.record()
            torch.cuda.synchronize(device)
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            t0 = time.perf_counter()
            one(o, d, gcol)
            dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
        if budget_s is not None and k >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    total = sum(times)
    what = "the reference's render_sh_voxel_grid + backward (oracle/_ref)" if kind == "reference" else "the oracle port of the reference"
    return {"rays_per_s": B * len(times) / total, "ms_per_step": 1e3 * total / len(times), "steps": len(times),
            "cores": 0 if on_gpu else threads, "kind": kind, "device": str(device),
            "sample": f"{len(times)} batches of {B} rays (pose 0, batches 17..) fwd+bwd through {what}, fp32, stock ATen ops, "
                      f"after {warmup} warm-up" + (", CUDA events" if on_gpu else f", {threads} host threads")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    device = torch.device(args.ref_device)
    r = reference_leg(device, args.steps, args.warmup, budget_s=150.0)  # K steps, or as many as fit into 2.5 minutes
    line = {
        "impl": "reference", "metric": "rays/s fwd+bwd, 160^3 SH-0 grid, 400x400 render", "value": r["rays_per_s"], "unit": "rays/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WL["name"], "step": f"one 4096-ray batch fwd+bwd on {'the GPU, stock ATen ops' if device.type == 'cuda' else 'the host CPU'} "
                                                   "(bounded sample of the frame)"},
        "cpu_baseline": {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["rays_per_s"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


FULL_AFFINITY = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None  # before any pinning


def reference_subprocess(device, steps, warmup, timeout_s=240):
    """Run the reference arm in a child process (the staged reference and this repository's mirror cannot share one
    interpreter) and return its JSON line, or a dict with "error"."""
    cmd = [sys.executable, str(Path(__file__).resolve()), "--impl", "reference", "--ref-device", device, "--steps", str(steps),
           "--warmup", str(warmup)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}

    def all_cpus():  # the child is the reference with every host thread it can use, whatever this process pinned itself to
        if FULL_AFFINITY and hasattr(os, "sched_setaffinity"):
            os.sched_setaffinity(0, FULL_AFFINITY)

    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env, cwd=str(ROOT), preexec_fn=all_cpus)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": (r.stderr or r.stdout)[-400:]}
        return json.loads(lines[-1])
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)[:400]}


# ---------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------
class DeviceBench:
    """Drives the C ABI directly (ctypes) on device-resident inputs; frames are captured as CUDA graphs."""

    N_GRID_COPIES = 3  # rotating copies: 3 x 65.5 MB of grid + 65.5 MB of gradients >> 126 MB L2

    def __init__(self, device, rank, world, count_s_in=True, n_lanes=1, kernel_jitter=True, postact=None, peer_volume=None,
                 share=None, sparse=False):
        from voxe_b200 import _native as nat
        from voxe_b200.render_function import FusedGridSpec, FusedRenderSpec, pack_volume

        self.nat, self.lib = nat, nat.load_library()
        self.device, self.rank, self.world = device, rank, world
        self.postact = postact or WL["postact"]
        dens, feat = (share.dens, share.feat) if share is not None else make_grid_tensors(device)
        half = [w / 2 for w in WL["world"]]
        self.N_GRID_COPIES = WL.get("grid_copies", self.N_GRID_COPIES)
        self.gspec = FusedGridSpec(dims=WL["dims"], n_features=feat.shape[-1], aabb=tuple((-h, h) for h in half),
                                   density_scale=WL["density_scale"], preact=nat.PREACT_IDENTITY,
                                   postact=nat.POSTACT_RELU if self.postact == "relu" else nat.POSTACT_SOFTPLUS)
        flags = nat.FLAG_WHITE_BKGD | (nat.FLAG_PERTURB if WL["perturb"] else 0)
        self.rspec = FusedRenderSpec(num_samples=WL["S"], near=WL["near"], far=WL["far"], flags=flags, sh_degree=WL["sh_degree"], n_colour=3)
        self.gd = self.gspec.to_native()
        self.rd = nat.VoxeRenderDesc.from_buffer_copy(self.rspec.native_bytes())  # private copy: rng_offset changes per launch
        self.kernel_jitter = kernel_jitter and WL["perturb"]
        self.dens, self.feat = dens, feat
        # a second bench on the same values (the Softplus twin) shares volumes, rays and outputs with the first
        self.packed = share.packed if share is not None else [pack_volume(self.gspec, dens, feat) for _ in range(self.N_GRID_COPIES)]
        if share is not None:
            self.packed_grad, self.peer_volume = share.packed_grad, share.peer_volume
        elif peer_volume is not None:  # N > 1: the gradient volume lives in peer-mapped memory (voxe_allreduce_grads_peer)
            self.peer_volume = peer_volume(self.packed[0].numel())
            self.packed_grad = self.peer_volume.buffer
        else:
            self.peer_volume = None
            self.packed_grad = torch.zeros_like(self.packed[0])
        self.d_dens, self.d_feat = (share.d_dens, share.d_feat) if share is not None else (torch.empty_like(dens), torch.empty_like(feat))
        # sparse hand-over: the backward leaves a trail of brick flags; the step's exchange (N > 1) and the conversion into
        # dense gradients follow it instead of sweeping the whole volume (voxe_allreduce_grads_peer_sparse, voxe_consume_grad)
        self.sparse, self.tag = bool(sparse), 1
        if share is not None:
            self.sparse, self.touched = share.sparse, share.touched
        elif not self.sparse:
            self.touched = None
        elif self.peer_volume is not None:
            self.touched = self.peer_volume.enable_sparse(self.gspec)
        else:
            self.touched = torch.zeros(int(self.lib.voxe_touched_bytes(self.gd)), dtype=torch.uint8, device=device)
        if WL.get("random_views"):
            # cfg 4 as BASELINE.json states it: `random_views` poses from get_random_pose after np.random.seed(42)
            # (utils/imaging_utils.py:197-215 upstream), dealt round-robin over the ranks
            from thre3d_atom.utils.imaging_utils import get_random_pose
            from voxe_b200.dist import shard_views

            np.random.seed(WL["seed"])
            every = [get_random_pose(WL["radius"])[0] for _ in range(WL["random_views"])]
            self.poses = [every[i] for i in shard_views(len(every), rank, world)]
            self.total_views = len(every)
        else:
            self.poses = make_poses()
        if WL["batch"] < WL["height"] * WL["width"] and WL["name"].startswith("cfg5"):
            self.poses = self.poses[:3]
        self.rays = share.rays if share is not None else [frame_rays(p, device) for p in self.poses]
        if WL["name"].startswith("cfg5"):  # one 65536-ray batch through the middle of each frame
            lo = (WL["height"] // 2 - 32) * WL["width"]
            self.rays = [(o[lo:lo + WL["batch"]].contiguous(), d[lo:lo + WL["batch"]].contiguous()) for o, d in self.rays]
        if WL.get("shuffle_rays") and share is None:
            # measurement variant: the frame's rays in random order, as the reconstruction loop's batches are
            # (sample_random_rays_and_pixels_synchronously, misc.py:126-138 upstream): no two neighbouring rays of a batch are
            # neighbouring pixels
            gp = torch.Generator(device=device).manual_seed(11)
            perms = [torch.randperm(o.shape[0], device=device, generator=gp) for o, _ in self.rays]
            self.rays = [(o[pm].contiguous(), d[pm].contiguous()) for (o, d), pm in zip(self.rays, perms)]
        self.R = self.rays[0][0].shape[0]
        g = torch.Generator().manual_seed(7)
        self.G = torch.randn(self.R, 3, generator=g).to(device)
        self.colour = torch.empty(self.R, 3, device=device)
        self.depth = torch.empty(self.R, device=device)
        self.acc = torch.empty(self.R, device=device)
        self.disp = torch.empty(self.R, device=device)
        self.n_lanes = max(1, n_lanes)
        self.jitter_ring = [torch.rand(WL["batch"], WL["S"], device=device) for _ in range(self.n_lanes + 1)]
        self.jitter = self.jitter_ring[0]
        self.side = torch.cuda.Stream(device)
        self.lane_streams = [torch.cuda.Stream(device) for _ in range(self.n_lanes - 1)]
        self.saved = torch.empty(int(self.lib.voxe_saved_floats(self.rd, WL["batch"])), device=device)
        self.saved_lane = [self.saved] + [torch.empty_like(self.saved) for _ in range(self.n_lanes - 1)]
        self.batches = [(s, min(s + WL["batch"], self.R)) for s in range(0, self.R, WL["batch"])]
        self.s_in = [count_inside_samples(o, d) for (o, d) in self.rays] if count_s_in else None
        # forward + backward per batch, the unpack, and (buffer mode only) one torch.rand per batch; the memset of the
        # gradient volume is not a kernel of this library
        self.kernels_per_step = 2 * len(self.batches) + 1

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _jitter_arg(self, pose, b0):
        """Explicit torch-drawn [R,S] jitter buffer, or None + a per-batch RNG offset for the in-kernel draws (the
        descriptor is read at launch time, so forward and backward of a batch see the same offset)."""
        if not WL["perturb"]:
            return None
        if self.kernel_jitter:
            self.rd.rng_seed, self.rd.rng_offset = WL["seed"], (pose << 20) | b0
            return None
        return self.jitter.data_ptr()

    def _fwd(self, pose, copy, b0, b1, saved):
        o, d = self.rays[pose]
        self.nat.check(self.lib.voxe_render_fwd(
            self.gd, self.rd, self.packed[copy].data_ptr(), o[b0:b1].data_ptr(), d[b0:b1].data_ptr(),
            self._jitter_arg(pose, b0), None, self.colour[b0:b1].data_ptr(), self.depth[b0:b1].data_ptr(),
            self.acc[b0:b1].data_ptr(), self.disp[b0:b1].data_ptr(), saved.data_ptr(), b1 - b0, self._stream()), "voxe_render_fwd")

    def _bwd(self, pose, copy, b0, b1, saved):
        o, d = self.rays[pose]
        self.nat.check(self.lib.voxe_render_bwd(
            self.gd, self.rd, self.packed[copy].data_ptr(), o[b0:b1].data_ptr(), d[b0:b1].data_ptr(),
            self._jitter_arg(pose, b0), None, saved.data_ptr(), self.G[b0:b1].data_ptr(), None, None, None,
            self.packed_grad.data_ptr(), self.touched.data_ptr() if self.sparse else None, self.tag if self.sparse else 0, b1 - b0,
            self._stream()), "voxe_render_bwd")

    def scatter_stats(self, pose=0):
        """(in-grid samples the backward processed, samples whose 8-corner scatter it issued) for one frame, counted by the
        kernel itself (VoxeRenderDesc.stats); run once, eagerly, outside every timed region."""
        counters = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.rd.stats = counters.data_ptr()
        n_bricks = int(self.lib.voxe_touched_bytes(self.gd))
        keep = (self.sparse, self.touched)
        if not self.sparse:  # a private trail, just to count the bricks a frame touches
            self.sparse, self.touched = True, torch.zeros(n_bricks, dtype=torch.uint8, device=self.device)
        else:
            self.touched.zero_()
        try:
            for b0, b1 in self.batches:
                self._fwd(pose, 0, b0, b1, self.saved)
                self._bwd(pose, 0, b0, b1, self.saved)
            torch.cuda.synchronize(self.device)
            self.touched_fraction = round(float((self.touched[:n_bricks] == self.tag).float().mean()), 4)
        finally:
            self.rd.stats = None
            self.sparse, self.touched = keep
        self.packed_grad.zero_()
        n_in, n_scatter, cell_leaders, corner_leaders = (int(v) for v in counters.tolist())
        # intra-warp duplicates among the scattering lanes of a warp instruction (8 neighbouring rays x 4 depth ranges)
        self.merge_stats = {"scattering_samples": n_scatter, "distinct_cells_per_warp_instruction": cell_leaders,
                            "corner_reds": 8 * n_scatter, "distinct_voxels_per_warp_instruction": corner_leaders,
                            "reds_after_perfect_intra_warp_merge": round(corner_leaders / max(1, 8 * n_scatter), 4)}
        return n_in, n_scatter

    def unpack(self):
        """Packed gradient volume -> dense d_densities / d_features (the reference's layout)."""
        if self.sparse:  # only the flagged bricks; what is read is zeroed, so the packed volume is all-zero again afterwards
            self.nat.check(self.lib.voxe_consume_grad(self.gd, self.packed_grad.data_ptr(), self.d_dens.data_ptr(), self.d_feat.data_ptr(),
                                                      self.touched.data_ptr(), self.tag, self._stream()), "voxe_consume_grad")
            return
        self.nat.check(self.lib.voxe_unpack_grad(self.gd, self.packed_grad.data_ptr(), self.d_dens.data_ptr(), self.d_feat.data_ptr(), 0,
                                                 self._stream()), "voxe_unpack_grad")

    def frame_body(self, pose, copy, what="both", saved_set=None, refresh_jitter=True, zero=True):
        """One frame.  ``saved_set``: per-batch workspaces (needed when fwd and bwd of a batch are not adjacent).

        Stream structure of the whole-frame step: the jitter draw of a later batch (torch.rand, sample.py:63) runs on a side
        stream while earlier batches render, and ``self.n_lanes`` batches are in flight at once, each lane running
        fwd(k) -> bwd(k) in order on its own stream with its own workspace.  Within a frame gradients are accumulated (one
        optimiser step per frame), so batch k+1's forward does not depend on batch k's backward; the scatter-adds of
        concurrent backward kernels into the shared gradient volume are atomic."""
        main = torch.cuda.current_stream(self.device)
        if what != "both":
            for k, (b0, b1) in enumerate(self.batches):
                saved = self.saved if saved_set is None else saved_set[k]
                if what == "fwd":
                    self._fwd(pose, copy, b0, b1, saved)
                else:
                    self._bwd(pose, copy, b0, b1, saved)
            return
        draw = WL["perturb"] and refresh_jitter and not self.kernel_jitter
        if zero and self.sparse:  # fresh dense gradients and a fresh trail; the packed volume was left all-zero by the consume
            self.d_dens.zero_()
            self.d_feat.zero_()
            self.touched.zero_()
        elif zero:  # one zero-fill per optimiser step (cfg 4 accumulates all of a rank's views before its one all-reduce)
            self.packed_grad.zero_()
        n = self.n_lanes
        lanes = [main] + self.lane_streams[: n - 1]
        for lane in lanes[1:]:
            lane.wait_stream(main)
        if draw:
            self.side.wait_stream(main)
        ring = len(self.jitter_ring)
        done = {}
        for k, (b0, b1) in enumerate(self.batches):
            lane = lanes[k % n]
            ready = None
            if draw:
                jit = self.jitter_ring[k % ring]
                with torch.cuda.stream(self.side):
                    if k >= ring:
                        self.side.wait_event(done[k - ring])
                    jit.uniform_()
                    ready = torch.cuda.Event()
                    ready.record(self.side)
                self.jitter = jit
            with torch.cuda.stream(lane):
                if ready is not None:
                    lane.wait_event(ready)
                saved = self.saved_lane[k % n] if saved_set is None else saved_set[k]
                self._fwd(pose, copy, b0, b1, saved)
                self._bwd(pose, copy, b0, b1, saved)
                done[k] = torch.cuda.Event()
                done[k].record(lane)
        for lane in lanes[1:]:
            main.wait_stream(lane)
        if draw:
            main.wait_stream(self.side)

    def capture(self, what="both", poses=None):
        graphs = []
        for n, pose in enumerate(poses if poses is not None else range(len(self.poses))):
            copy = n % self.N_GRID_COPIES
            saved_set = None
            if what != "both":  # isolated kernels: fixed jitter, per-batch workspaces filled by a full pass first
                saved_set = [torch.empty_like(self.saved) for _ in self.batches]
                self.frame_body(pose, copy, "both", saved_set, refresh_jitter=False)
            zero = n == 0 or not WL.get("random_views")
            self.frame_body(pose, copy, what, saved_set, zero=zero)  # warm (lazy module load, allocator)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.frame_body(pose, copy, what, saved_set, zero=zero)
            graphs.append(g)
            self._keep = getattr(self, "_keep", []) + [saved_set]
        return graphs

    def time_graphs(self, graphs, steps, warmup, after_replay=None, barrier=None):
        n = len(graphs)
        for k in range(warmup):
            graphs[(k * self.world + self.rank) % n].replay()
            if after_replay:
                after_replay()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if barrier:
            barrier()
        torch.cuda.synchronize(self.device)
        start.record()
        for k in range(steps):
            graphs[((warmup + k) * self.world + self.rank) % n].replay()
            if after_replay:
                after_replay()
        stop.record()
        torch.cuda.synchronize(self.device)
        if barrier:
            barrier()
        return start.elapsed_time(stop)  # ms


def e2e_leg(device, rank, world, steps, warmup, dist, deferred=False, engine_threads=True, whole_frame=False, graph=False,
            peer_volume=None, frame_backward=False):
    """The frame through the public API, inputs starting in pinned host memory.

    engine_threads  True: torch's default autograd engine (a device worker thread runs every backward()); False: the stock
                    caller-side switch torch.autograd.set_multithreading_enabled(False) (backward on the calling thread)
    graph           the same API calls of one frame captured ONCE with torch.cuda.graph (stock torch; the render path is
                    capturable: in-kernel jitter reads the graph-registered generator state, so every replay draws fresh
                    jitter) and replayed per frame on static device buffers that the per-frame host->device copies fill
    deferred        the opt-in gradient accumulation mode (VoxelGrid.accumulate_render_gradients), one materialisation per frame
    whole_frame     one render_rays + one backward per frame instead of 4096-ray batches
    N > 1: one collective per frame -- on the packed sink volume (deferred; the library's peer kernel when the volume is
    peer-mapped) or on ONE flat buffer holding both dense gradients (VoxelGradAllReducer)."""
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.render_interface import Rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds
    from voxe_b200.dist import VoxelGradAllReducer

    dens, feat = make_grid_tensors(device)
    grid = VoxelGrid(dens, feat, VoxelSize(*(w / d for w, d in zip(WL["world"], WL["dims"]))), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=WL["density_scale"], tunable=True)
    vm = VolumetricModel(grid, render_sh_voxel_grid,
                         SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]),
                                               white_bkgd=True, perturb_sampled_points=WL["perturb"]), device=device)
    volume = None
    if deferred:
        grid.accumulate_render_gradients()
        if world > 1 and peer_volume is not None:
            spec = grid.fused_spec()
            packed = grid.packed_cache().get(spec, grid.densities, grid.features)
            volume = peer_volume(packed.numel())
            volume.adopt(grid.render_gradient_accumulator)
    # N > 1, dense gradients: .grad of both parameters are views into ONE peer-mapped buffer reduced in place by the library's
    # kernel (PeerGradients); without peer mapping, one flat staging buffer through ncclAllReduce (VoxelGradAllReducer)
    reducer = peer_grads = None
    if world > 1 and not deferred:
        if peer_volume is not None:
            from voxe_b200.dist import PeerGradients

            peer_grads = PeerGradients([grid.densities, grid.features])
        else:
            reducer = VoxelGradAllReducer([grid.densities, grid.features])

    def clear_grads():
        if peer_grads is not None:
            peer_grads.zero()  # in place: the views stay attached (optimizer.zero_grad(set_to_none=False) semantics)
        else:
            grid.densities.grad = None
            grid.features.grad = None
    poses = make_poses()
    host = []
    g = torch.Generator().manual_seed(7)
    for p in poses:
        o, d = frame_rays(p, torch.device("cpu"))
        host.append((o.pin_memory(), d.pin_memory(), torch.randn(o.shape[0], 3, generator=g).pin_memory()))
    R, B = host[0][0].shape[0], WL["batch"]
    if whole_frame:
        B = R  # one render_rays + one backward per frame (how the SDS edit loop calls it, sds_trainer.py:283)
    colour_host = torch.empty(R, 3).pin_memory()
    h2d = R * (12 + 12 + 12)
    d2h = R * 12 + 4

    def render_frame(o, d, gc):
        """The API calls of one frame on device-resident inputs: per batch render_rays -> backward."""
        colours, roots = [], []
        g_batches = gc.split(B)
        for o_b, d_b, g_b in zip(o.split(B), d.split(B), g_batches):  # 4096-ray batches (views, no copies)
            out = vm.render_rays(Rays(o_b, d_b))
            # loss = <colour, G>: the upstream gradient is handed to autograd directly, as the SDS step does
            # (thre3d_reprs/sd.py:20-34 SpecifyGradient)
            if frame_backward:
                roots.append(out.colour)
            else:
                out.colour.backward(g_b)
            colours.append(out.colour.detach())
        if frame_backward:  # one engine invocation for the frame's 40 render nodes: the loss of a frame is the sum over its batches
            torch.autograd.backward(roots, grad_tensors=list(g_batches))
        colour = torch.cat(colours)
        return colour, (colour * gc).sum()

    def finish(colour, loss_total):
        colour_host.copy_(colour, non_blocking=True)  # the step's result: the rendered frame and the loss
        if deferred:
            if volume is not None:
                volume.allreduce()  # ONE collective on the packed volume: the library's own peer-memory kernel
                grid.render_gradient_accumulator.dirty = True
            elif world > 1:
                dist.all_reduce(grid.render_gradient_accumulator.buffer)
            grid.materialize_render_gradients()
        elif peer_grads is not None:
            peer_grads.allreduce()  # ONE collective, in place on the memory .grad lives in
        elif world > 1:
            reducer()  # ONE collective: both dense gradients through one flat staging buffer
        return float(loss_total.item())  # D2H of the step's result; also orders the colour copy

    static = None
    if graph:
        static = [torch.empty(R, 3, device=device) for _ in range(3)]
        for t, h in zip(static, host[0]):
            t.copy_(h)
        clear_grads()
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):  # warm-up on a side stream, as torch's CUDA-graph recipe asks
            for _ in range(2):
                clear_grads()
                render_frame(*static)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        clear_grads()
        cuda_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cuda_graph):
            if peer_grads is not None:
                peer_grads.zero()  # part of the replayed frame
            static_colour, static_loss = render_frame(*static)

    # Input pipeline: the host -> device copy of a step's inputs (rays + upstream gradients of the whole frame, pinned
    # memory) is issued on a copy stream one step ahead, the way a training loop's data loader prefetches; every step's
    # copy still happens inside the timed region, overlapped with the previous step's kernels.
    copy_stream = torch.cuda.Stream(device)
    staged = {}

    def stage(idx):
        with torch.cuda.stream(copy_stream):
            bufs = [h.to(device, non_blocking=True) for h in host[idx % len(host)]]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[idx] = (bufs, ev)

    def one_frame(idx, nxt=None):
        if idx not in staged:
            stage(idx)
        (o, d, gc), ev = staged.pop(idx)
        if nxt is not None:
            stage(nxt)
        torch.cuda.current_stream(device).wait_event(ev)
        if graph:
            # the step's inputs land in the graph's static buffers (device -> device); .grad is re-zeroed by the graph itself
            # (the first backward of the captured frame created it with a zero-fill that is part of the graph)
            for t, src in zip(static, (o, d, gc)):
                t.copy_(src, non_blocking=True)
            cuda_graph.replay()
            return finish(static_colour, static_loss)
        clear_grads()
        return finish(*render_frame(o, d, gc))

    with torch.autograd.set_multithreading_enabled(engine_threads):
        for k in range(warmup):
            one_frame(k * world + rank)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        frames = [(warmup + k) * world + rank for k in range(steps)]
        t0 = time.perf_counter()
        for k, idx in enumerate(frames):
            one_frame(idx, frames[k + 1] if k + 1 < steps else None)
        torch.cuda.synchronize(device)
        elapsed = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([elapsed], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    api = ("per 4096-ray batch: VolumetricModel.render_rays(Rays) -> out.colour.backward(dL/dcolour); per frame: one pinned H2D copy "
           "of rays + upstream gradients (issued one frame ahead on a copy stream, inside the timed region), one D2H copy of the "
           "rendered colours and the loss")
    if graph:
        api += "; the frame's API calls captured once with torch.cuda.graph and replayed per frame on static input buffers"
    return {"value": world * R * steps / elapsed, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": 1e3 * elapsed / steps, "steps": steps,
            "autograd_engine": "worker threads (torch default)" if engine_threads else "calling thread (torch.autograd.set_multithreading_enabled(False))",
            "api": api}


def inference_leg(device, frames=24):
    """Rows f3/f4: ``VolumetricModel.render`` of the benchmark camera under no_grad.  The whole-camera kernel (rays
    generated in-kernel, early termination at T < 1e-5) against the reference's structure of the same call -- cast_rays
    tensors walked in 32768-ray chunks (modules/volumetric_model.py:170-186) through the training kernels."""
    from thre3d_atom.modules.volumetric_model import VolumetricModel
    from thre3d_atom.rendering.volumetric.utils.misc import cast_rays, flatten_rays
    from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from thre3d_atom.utils.imaging_utils import CameraBounds, CameraIntrinsics

    dens, feat = make_grid_tensors(device)
    grid = VoxelGrid(dens, feat, VoxelSize(*(w / d for w, d in zip(WL["world"], WL["dims"]))), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=WL["density_scale"], tunable=False)
    vm = VolumetricModel(grid, render_sh_voxel_grid,
                         SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]),
                                               white_bkgd=True, perturb_sampled_points=False), device=device)
    cam = CameraIntrinsics(WL["height"], WL["width"], WL["focal"])
    poses = make_poses()

    def chunked(pose):
        rays = flatten_rays(cast_rays(cam, pose, device=device))
        with torch.no_grad():
            return [vm.render_rays(rays[s:s + 32768]).colour for s in range(0, len(rays), 32768)]

    def timed(fn):
        for k in range(3):
            fn(poses[k % len(poses)])
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for k in range(frames):
            fn(poses[k % len(poses)])
        torch.cuda.synchronize(device)
        return (time.perf_counter() - t0) / frames

    t_cam = timed(lambda pose: vm.render(pose, cam))
    t_chunk = timed(chunked)
    rays = WL["height"] * WL["width"]
    return {"api": "VolumetricModel.render (no_grad), 400x400, S=256", "ms_per_frame": round(1e3 * t_cam, 3), "rays_per_s": rays / t_cam,
            "chunked_ray_tensor_route": {"ms_per_frame": round(1e3 * t_chunk, 3), "rays_per_s": rays / t_chunk,
                                         "note": "cast_rays + 32768-ray chunks through the training kernels (the structure of the reference's render())"}}


def fused_step_leg(device, peak):
    """Row f2: the optimiser-step grid pass.  Fused kernel (consume packed grads + Adam + repack + zero) against the
    unfused sequence the reference-style loop runs (zero-fill, unpack, torch.optim.Adam.step, repack), CUDA events."""
    from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
    from voxe_b200.optim import FusedVoxelAdam
    from voxe_b200.render_function import pack_volume

    def timed(fn, prepare, n=20, warm=3):
        """Mean device time of fn(): [prepare(); fn()] x n back to back minus [prepare()] x n, CUDA events on the stream.
        prepare() restores what a backward pass leaves behind (a non-trivial gradient volume)."""
        def loop(with_fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(device)
            a.record()
            for _ in range(n):
                prepare()
                if with_fn:
                    fn()
            b.record()
            torch.cuda.synchronize(device)
            return a.elapsed_time(b)
        for _ in range(warm):
            prepare()
            fn()
        return 1e3 * (loop(True) - loop(False)) / n  # us

    dens, feat = make_grid_tensors(device)
    grid = VoxelGrid(dens, feat, VoxelSize(*(w / d for w, d in zip(WL["world"], WL["dims"]))), density_preactivation=torch.nn.Identity(),
                     density_postactivation=torch.nn.ReLU(), expected_density_scale=WL["density_scale"], tunable=True)
    spec = grid.fused_spec()
    packed = grid.packed_cache().get(spec, grid.densities, grid.features)
    opt = FusedVoxelAdam(grid, lr=0.03)
    acc = grid.render_gradient_accumulator
    buf = acc.get(packed)
    noise = torch.randn_like(packed)

    def scattered():  # what a backward pass leaves behind: a non-trivial gradient volume
        buf.copy_(noise)
        acc.dirty = True

    fused_us = timed(opt.step, scattered)
    channels = packed.numel()
    fused_bytes = 36.0 * dens.numel() * (feat.shape[-1] + 1)  # r: g,p,m,v  w: p,m,v,packed,zeroed g  (4 B each)

    ref_d, ref_f = torch.nn.Parameter(dens.clone()), torch.nn.Parameter(feat.clone())
    ref = torch.optim.Adam([{"params": [ref_d, ref_f], "lr": 0.03}], betas=(0.9, 0.999))
    pg = torch.empty_like(packed)
    ref_d.grad, ref_f.grad = torch.zeros_like(ref_d), torch.zeros_like(ref_f)
    lib, gd = grid_lib = (__import__("voxe_b200._native", fromlist=["x"]).load_library(), spec.to_native())
    stream = torch.cuda.current_stream(device).cuda_stream

    def unfused():
        lib.voxe_unpack_grad(gd, pg.data_ptr(), ref_d.grad.data_ptr(), ref_f.grad.data_ptr(), 1, stream)  # .grad += render grads
        pg.zero_()
        ref.step()
        pack_volume(spec, ref_d, ref_f, out=packed)
        ref_d.grad.zero_(), ref_f.grad.zero_()

    unfused_us = timed(unfused, lambda: pg.copy_(noise))
    return {"kernel": "adam_step_kernel (voxe_adam_step)", "us": round(fused_us, 1), "bytes": int(fused_bytes),
            "achieved": round(fused_bytes / (fused_us * 1e-6) / 1e9, 1), "unit": "GB/s", "frac": round(fused_bytes / (fused_us * 1e-6) / 1e9 / peak, 4),
            "unfused_us": round(unfused_us, 1), "unfused": "unpack(+=) + zero-fill + torch.optim.Adam.step + repack + grad zero", "packed_floats": channels}


def regularizers_leg(device, peak):
    """Row f2: the per-step regularisers of the edit loop on the benchmark grid -- density-correlation loss (weight 200 by
    default in the reference's edit script) and total variation of ReLU(densities) and of the features -- loss + gradient
    accumulated into .grad, against the same formulas as torch ops on the same GPU (the reference's code path:
    sds_trainer.py:290-326 + autograd).  CUDA events, mean of 20."""
    from voxe_b200 import regularizers as reg

    def timed(fn, n=20, warm=3, graph=False):
        """Mean time of fn() in us, CUDA events.  graph=True: n calls captured into one CUDA graph and replayed, i.e. the
        kernels back to back without the Python / ctypes time of the call (which exceeds the kernels' at this grid size)."""
        for _ in range(warm):
            fn()
        run = lambda: [fn() for _ in range(n)]  # noqa: E731
        if graph:
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side), torch.cuda.graph(g, stream=side):
                run()
            torch.cuda.current_stream(device).wait_stream(side)
            run = g.replay
            run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        a.record()
        run()
        b.record()
        torch.cuda.synchronize(device)
        return 1e3 * a.elapsed_time(b) / n  # us

    dens, feat = make_grid_tensors(device)
    pre = (dens + 0.1 * torch.randn_like(dens)).contiguous()
    dens, feat = torch.nn.Parameter(dens), torch.nn.Parameter(feat)
    dens.grad, feat.grad = torch.zeros_like(dens), torch.zeros_like(feat)

    def tv_torch(g):
        return (g.diff(dim=0).abs().mean() + g.diff(dim=1).abs().mean() + g.diff(dim=2).abs().mean()) / 3

    def corr_torch(a, b):  # sds_trainer.py:507-524
        cov = (a - torch.mean(a)) * (b - torch.mean(b))
        den = torch.sqrt(torch.mean((a - torch.mean(a)) ** 2) * torch.mean((b - torch.mean(b)) ** 2))
        return 1.0 - torch.mean(cov / (den + 1e-7))

    out = {}
    cases = [
        ("density_correlation", lambda: reg.accumulate_density_loss_gradient(dens, pre, 200.0),
         lambda: (corr_torch(dens, pre) * 200.0).backward(), 4.0 * dens.numel() * (2 + 2 + 2)),  # stats: a,b; grad: a,b + .grad rmw
        ("tv_relu_densities", lambda: reg.accumulate_tv_gradient(dens, 1.0, relu=True),
         lambda: tv_torch(torch.relu(dens)).backward(), 4.0 * dens.numel() * 3),  # grid read + .grad rmw
        ("tv_features", lambda: reg.accumulate_tv_gradient(feat, 1.0),
         lambda: tv_torch(feat).backward(), 4.0 * feat.numel() * 3),
    ]
    for name, ours, theirs, nbytes in cases:
        call_us = timed(ours)
        us = timed(ours, graph=True)
        torch_us = timed(theirs, n=5, warm=2)
        out[name] = {"us": round(us, 1), "call_us": round(call_us, 1), "torch_ops_us": round(torch_us, 1), "bytes": int(nbytes),
                     "achieved": round(nbytes / (us * 1e-6) / 1e9, 1), "unit": "GB/s", "frac": round(nbytes / (us * 1e-6) / 1e9 / peak, 4)}
    out["note"] = ("loss + weight * gradient accumulated into .grad (voxe_pair_loss + voxe_pair_loss_grad / voxe_tv_regularizer); "
                   "us = kernels back to back (graph replay), call_us = through the Python call; "
                   "torch_ops_us = the reference's formulas as torch ops + autograd on the same GPU")
    return out


def sampler_leg(device):
    """Row f3: one training batch drawn from 8 views of 800x800 (the reference loop's shape, trainers.py:290-313): 4096
    distinct pixels, their rays and colours.  Ours: one launch from poses (no ray tensors).  torch_ops_us: the reference's
    per-iteration code on the same GPU -- randperm over all 5.12 M pixels and three gathers from pre-cast rays (the casting
    itself, done once per loaded batch of views upstream, is not counted)."""
    from thre3d_atom.utils.imaging_utils import CameraIntrinsics
    from voxe_b200 import sampling

    b, h, w, k = 8, 800, 800, 4096
    n = b * h * w
    intr = CameraIntrinsics(h, w, 1111.1)
    poses = torch.eye(3, 4, device=device).repeat(b, 1, 1).contiguous()
    pixels = torch.rand(n, 3, device=device)
    rays_o, rays_d = torch.rand(n, 3, device=device), torch.rand(n, 3, device=device)

    def timed(fn, n_it=50, warm=5):
        for _ in range(warm):
            fn()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        a.record()
        for _ in range(n_it):
            fn()
        e.record()
        torch.cuda.synchronize(device)
        return 1e3 * a.elapsed_time(e) / n_it

    def reference_ops():
        chosen = torch.randperm(n, dtype=torch.long, device=device)[:k]
        return rays_o[chosen, :], rays_d[chosen, :], pixels[chosen, :]

    return {"us": round(timed(lambda: sampling.sample_rays_from_cameras(intr, poses, pixels, k)), 1),
            "torch_ops_us": round(timed(reference_ops, n_it=10, warm=2), 1),
            "shape": f"{k} of {b} x {h} x {w} pixels",
            "note": "us = Python call to voxe_sample_rays (host-bound: the kernel is ~3 us); pre-cast rays avoided: "
                    f"{n * 24 / 1e6:.0f} MB"}


def oracle_frame(bench, pose, want_grads=True, ray_stride=1):
    """The oracle (stock ATen ops, fp32) on the bench's own GPU for frame ``pose`` with the very jitter the kernels drew
    (voxe_jitter_fill with the per-batch (seed, offset) of DeviceBench._jitter_arg).  Test infrastructure used as the
    checker of the timed leg, never timed as the product.  Returns colour [R,3] (NaN rows where ray_stride skipped) and,
    with want_grads, the frame's dense voxel gradients."""
    from oracle.voxe_oracle import OracleConfig, OracleGrid, relu_kink_voxels, render_oracle, render_oracle_with_grads

    o, d = bench.rays[pose]
    voxel = tuple(w / n for w, n in zip(WL["world"], WL["dims"]))
    ogrid = OracleGrid(voxel, density_scale=WL["density_scale"], preact=WL["preact"], postact=bench.postact)
    ocfg = OracleConfig(num_samples=WL["S"], near=WL["near"], far=WL["far"], perturb=bool(WL["perturb"]), white_bkgd=True)
    colour = torch.full((bench.R, 3), float("nan"), device=bench.device)
    gd = torch.zeros_like(bench.dens) if want_grads else None
    gf = torch.zeros_like(bench.feat) if want_grads else None
    # ReLU field: voxels that are corners of a sample whose interpolated density is within rounding of the kink (its
    # derivative is decided by fp32 evaluation order; tests/test_cuda_parity.py masks the same set)
    kink = torch.zeros(bench.dens.shape[:3], dtype=torch.bool, device=bench.device) if (want_grads and bench.postact == "relu") else None
    rd = bench.nat.VoxeRenderDesc.from_buffer_copy(bench.rspec.native_bytes())
    chunk = 8192
    for b0, b1 in bench.batches:
        jit = None
        if WL["perturb"]:
            if bench.kernel_jitter:
                rd.rng_seed, rd.rng_offset = WL["seed"], (pose << 20) | b0
                jit = torch.empty(b1 - b0, WL["S"], device=bench.device)
                bench.nat.check(bench.lib.voxe_jitter_fill(rd, jit.data_ptr(), b1 - b0, bench._stream()), "voxe_jitter_fill")
            else:
                jit = bench.jitter
        for c0 in range(b0, b1, chunk):
            c1 = min(c0 + chunk, b1)
            sel = torch.arange(c0, c1, ray_stride, device=bench.device)
            j = None if jit is None else jit[sel - b0]
            if want_grads:
                res = render_oracle_with_grads(bench.dens, bench.feat, ogrid, o[sel], d[sel], ocfg, bench.G[sel], jitter=j, dtype=torch.float32)
                gd += res["d_densities"]
                gf += res["d_features"]
                if kink is not None:
                    kink |= relu_kink_voxels(bench.dens, ogrid, o[sel], d[sel], ocfg, jitter=j)
            else:
                with torch.no_grad():
                    res = render_oracle(bench.dens, bench.feat, ogrid, o[sel], d[sel], ocfg, jitter=j, dtype=torch.float32)
            colour[sel] = res["colour"]
    return colour, gd, gf, kink


def parity_of_timed_leg(bench, pose, copy, world):
    """"the timed run is the same run that is parity-checked".  Called right after the timed loop: the
    bench's output buffers still hold the colours of the last timed frame (``pose``) and -- at N = 1 -- d_densities /
    d_features hold its unpacked voxel gradients.  Both are compared with the oracle on the same GPU (fp32, same jitter)."""
    colour_timed = bench.colour.clone()
    want_grads = world == 1
    got_d, got_f = (bench.d_dens.clone(), bench.d_feat.clone()) if want_grads else (None, None)
    colour, gd, gf, kink = oracle_frame(bench, pose, want_grads=want_grads)
    out = {"frame": f"last timed frame (pose {pose}, packed-volume copy {copy}), all {bench.R} rays",
           "oracle": "oracle/voxe_oracle.py, fp32 ATen ops on the same GPU, jitter = voxe_jitter_fill of the launches' (seed, offset)",
           "colour_max_abs": float((colour_timed - colour).abs().max()), "colour_tol": 1e-4}
    ok = out["colour_max_abs"] <= out["colour_tol"]
    if want_grads:
        # ReLU: a sample whose interpolated density is within rounding of 0 flips its derivative between two fp32
        # evaluation orders, which moves only that sample's scatter into its 8 corner voxels of d_densities: those voxels
        # (oracle.relu_kink_voxels, margin 2e-3) are excluded there, everything else is held to the tolerance
        tol = 2e-4
        for name, got, want in (("d_densities", got_d, gd), ("d_features", got_f, gf)):
            diff = (got - want)
            if kink is not None and name == "d_densities":
                diff = diff * (~kink)[..., None]
                out["relu_kink_voxels_excluded"] = {"count": int(kink.sum()), "of": kink.numel()}
            out[name] = {"rel_l2": float(diff.norm() / want.norm().clamp_min(1e-30)),
                         "max_abs_over_inf": float(diff.abs().max() / want.abs().max().clamp_min(1e-30))}
            ok = ok and out[name]["rel_l2"] <= tol and out[name]["max_abs_over_inf"] <= tol
        out["grad_tol"] = tol
    else:
        out["gradients"] = "N > 1: the unpacked volume is the sum over ranks; the rank-sum check is `bench.py --check`"
    out["ok"] = bool(ok)
    return out


def l2_probe(device):
    """Measured L2 bandwidth: an elementwise kernel (torch.mul) streaming one buffer into another, both resident in the
    126 MB L2 after the first pass; read + write bytes / time, best over buffer sizes of 8..40 MB and 20 runs of 10
    back-to-back launches each, CUDA events.  (Stock-torch probe: small sizes are launch-latency bound, large ones spill
    out of the L2, a same-dtype copy_ goes through the copy engine -- the best of the sizes is the figure reported.)"""
    best_gbs, best_mb = 0.0, 0
    for mb in (8, 16, 24, 32, 40):
        n = mb * 2**18
        a, b = torch.rand(n, device=device), torch.empty(n, device=device)
        for _ in range(5):
            torch.mul(a, 1.0001, out=b)
        best = float("inf")
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                torch.mul(a, 1.0001, out=b)
            e1.record()
            torch.cuda.synchronize(device)
            best = min(best, e0.elapsed_time(e1) / 10)
        gbs = 2 * n * 4 / (best * 1e-3) / 1e9
        if gbs > best_gbs:
            best_gbs, best_mb = gbs, mb
    return best_gbs, best_mb


def ncu_capture(kernel_substring):
    """Counters of a kernel from the committed ncu capture (profiles/r2_ncu_kernels.json, written by tools/ncu_extract.py
    from an `ncu --set full` run of `bench.py --ncu`; keyed by the git revision it was taken on)."""
    path = ROOT / "profiles" / "r2_ncu_kernels.json"
    if not path.exists():
        return None, None
    data = json.loads(path.read_text())
    for name, rec in data.get("kernels", {}).items():
        if kernel_substring in name:
            return rec, {"file": "profiles/r2_ncu_kernels.json", "git": data.get("git"), "command": data.get("command")}
    return None, None


def make_peer_volume_factory(device, world, collective):
    """N > 1: gradient volumes come from torch's symmetric-memory allocator (plumbing) so that the library's own all-reduce
    kernel can reach every rank's copy; None selects ncclAllReduce through torch.distributed."""
    if world == 1 or collective == "nccl":
        return None
    from voxe_b200.dist import PeerGradVolume

    # measured, 68 MB (profiles/r2_check_n*.json): plain peer loads / stores 131 / 181 / 210-214 us at N = 2 / 4 / 8, the
    # switch's multicast reduction 210 / 196 / 208 us -- the plain path up to four ranks, multimem above
    multicast = collective == "peer" and world > 4
    return lambda n_floats: PeerGradVolume(n_floats, device, multicast=multicast)


def red_issue_floor(merge_stats, launches_per_frame, kernel_us, clocks):
    if not merge_stats:
        return None
    lane_reds = merge_stats["corner_reds"] / launches_per_frame
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    floor_us = lane_reds * 1.29 / (148 * sm_mhz)
    return {"lane_reds_per_launch": round(lane_reds), "clk_per_lane_red": 1.29, "floor_us": round(floor_us, 2), "frac": round(floor_us / kernel_us, 4),
            "note": "share of the kernel's time the SM-side RED issue rate alone accounts for (148 SMs, measured SM clock)"}


def run_check(args):
    """`--check`: the CUDA path's gradients under data parallelism.  Every rank renders a disjoint, round-robin share of the
    4096-ray batches of one frame; the packed gradient volumes are summed with (a) the library's peer-memory kernel
    (multicast and plain peer loads), (b) ncclAllReduce through the C ABI on a communicator built with the voxe_nccl_*
    helpers, (c) torch.distributed; each sum is compared, on every rank, with the volume of an unsharded render of the
    whole frame on that rank (<= 1e-5 ||g||inf: float atomics order).  Prints one JSON line (rank 0)."""
    import torch.distributed as dist

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    WL["perturb"] = True
    bench = DeviceBench(device, rank, world, count_s_in=False, n_lanes=1, kernel_jitter=True)
    pose = 2

    def render(batches):
        bench.packed_grad.zero_()
        for b0, b1 in batches:
            bench._fwd(pose, 0, b0, b1, bench.saved)
            bench._bwd(pose, 0, b0, b1, bench.saved)
        torch.cuda.synchronize(device)
        return bench.packed_grad.clone()

    full = render(bench.batches)
    share = render(bench.batches[rank::world])
    scale = float(full.abs().max())
    results = {}

    def record(name, summed):
        err = float((summed - full).abs().max()) / scale
        t = torch.tensor([err], device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[name] = {"max_abs_over_inf": float(t.item()), "ok": float(t.item()) <= 1e-5}

    from voxe_b200 import _native as nat
    from voxe_b200.dist import PeerGradVolume

    lib = nat.load_library()
    if world > 1:
        variants = [("voxe_allreduce_grads_peer (multimem)", True, 8), ("voxe_allreduce_grads_peer (peer loads/stores)", False, 0)]
        if args.hybrid:  # tuning: split the CTAs of a launch between the two paths (measured at N = 8: no gain, 209-211 us)
            variants += [(f"voxe_allreduce_grads_peer (hybrid: {k} of 8 CTAs multimem)", True, k) for k in (6, 4)]
        for name, multicast, mc_share in variants:
            try:
                vol = PeerGradVolume(full.numel(), device, multicast=multicast, multicast_share=mc_share or None)
                if multicast and not vol.multicast:
                    results[name] = {"skipped": "no multicast mapping on this box"}
                    continue
                for _ in range(3):  # repeated use of the same signal pads
                    vol.buffer.copy_(share)
                    torch.cuda.synchronize(device)
                    dist.barrier()
                    vol.allreduce()
                    torch.cuda.synchronize(device)
                assert not vol.failed(), "a peer did not arrive"
                record(name, vol.buffer)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                torch.cuda.synchronize(device)
                e0.record()
                for _ in range(10):
                    vol.allreduce()
                e1.record()
                torch.cuda.synchronize(device)
                results[name]["us"] = round(1e3 * e0.elapsed_time(e1) / 10, 1)
                results[name]["bytes"] = full.numel() * 4
                if rank == 0:
                    print(f"[check] {name}: {results[name]}", file=sys.stderr, flush=True)
                del vol
            except Exception as exc:  # noqa: BLE001
                results[name] = {"error": str(exc)[:300]}
        # (a') the brick-wise exchange: a PART of the frame (so most bricks stay untouched), each rank's share rendered with
        # brick flags straight into the peer volume; three rounds with fresh tags over flags that are never cleared; then the
        # flag-guided hand-over into dense gradients, against the unpacked volume of the unsharded render
        part = bench.batches[: max(world, len(bench.batches) // 6)]
        full_part = render(part)
        want_d, want_f = torch.empty_like(bench.d_dens), torch.empty_like(bench.d_feat)
        nat.check(lib.voxe_unpack_grad(bench.gd, full_part.data_ptr(), want_d.data_ptr(), want_f.data_ptr(), 0, bench._stream()), "voxe_unpack_grad")
        for name, multicast in (("voxe_allreduce_grads_peer_sparse (multimem)", True), ("voxe_allreduce_grads_peer_sparse (peer loads/stores)", False)):
            keep = (bench.packed_grad, bench.sparse, bench.touched, bench.tag)
            try:
                vol = PeerGradVolume(full.numel(), device, multicast=multicast)
                if multicast and not vol.multicast:
                    results[name] = {"skipped": "no multicast mapping on this box"}
                    continue
                bench.packed_grad, bench.sparse, bench.touched = vol.buffer, True, vol.enable_sparse(bench.gspec)
                for tag in (1, 2, 3):
                    bench.tag = tag
                    render(part[rank::world])  # zero-fills the volume first; leaves this rank's share and its flags
                    dist.barrier()
                    vol.allreduce_sparse(tag)
                    torch.cuda.synchronize(device)
                assert not vol.failed(), "a peer did not arrive"
                err = float((vol.buffer - full_part).abs().max()) / scale
                n_bricks = int(lib.voxe_touched_bytes(bench.gd))
                frac = float((bench.touched[:n_bricks] == bench.tag).float().mean())
                bench.d_dens.zero_()
                bench.d_feat.zero_()
                bench.unpack()  # voxe_consume_grad along the union of the flags
                torch.cuda.synchronize(device)
                err = max(err, float((bench.d_dens - want_d).abs().max()) / scale, float((bench.d_feat - want_f).abs().max()) / scale)
                left = float(vol.buffer.abs().max())
                t = torch.tensor([err, left], device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                results[name] = {"max_abs_over_inf": float(t[0]), "volume_left_after_hand_over": float(t[1]), "ok": float(t[0]) <= 1e-5 and float(t[1]) == 0.0,
                                 "flagged_brick_fraction_after_union": round(frac, 4), "frame_part": f"{len(part)} of {len(bench.batches)} batches"}
                render(part[rank::world])
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                torch.cuda.synchronize(device)
                e0.record()
                for _ in range(10):
                    vol.allreduce_sparse(bench.tag)
                e1.record()
                torch.cuda.synchronize(device)
                results[name]["us"] = round(1e3 * e0.elapsed_time(e1) / 10, 1)
                if rank == 0:
                    print(f"[check] {name}: {results[name]}", file=sys.stderr, flush=True)
                del vol
            except Exception as exc:  # noqa: BLE001
                results[name] = {"error": str(exc)[:300]}
            finally:
                bench.packed_grad, bench.sparse, bench.touched, bench.tag = keep
        # (b) the C ABI's NCCL entry point on its own communicator
        try:
            import ctypes

            uid = torch.zeros(nat.NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
            if rank == 0:
                nat.check(lib.voxe_nccl_unique_id(uid.data_ptr()), "voxe_nccl_unique_id")
            uid_dev = uid.to(device)
            dist.broadcast(uid_dev, 0)
            uid = uid_dev.cpu()
            comm = ctypes.c_void_p()
            nat.check(lib.voxe_nccl_comm_create(ctypes.byref(comm), world, rank, uid.data_ptr()), "voxe_nccl_comm_create")
            buf = share.clone()
            stream = torch.cuda.current_stream(device).cuda_stream
            nat.check(lib.voxe_allreduce_grads(comm, buf.data_ptr(), buf.numel(), stream), "voxe_allreduce_grads")
            torch.cuda.synchronize(device)
            record("voxe_allreduce_grads (ncclAllReduce, own communicator)", buf)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            torch.cuda.synchronize(device)
            e0.record()
            for _ in range(10):
                lib.voxe_allreduce_grads(comm, buf.data_ptr(), buf.numel(), stream)
            e1.record()
            torch.cuda.synchronize(device)
            results["voxe_allreduce_grads (ncclAllReduce, own communicator)"]["us"] = round(1e3 * e0.elapsed_time(e1) / 10, 1)
            nat.check(lib.voxe_nccl_comm_destroy(comm), "voxe_nccl_comm_destroy")
        except Exception as exc:  # noqa: BLE001
            results["voxe_allreduce_grads (ncclAllReduce, own communicator)"] = {"error": str(exc)[:300]}
        buf = share.clone()
        dist.all_reduce(buf)
        record("torch.distributed.all_reduce (NCCL)", buf)
    else:
        record("single rank (no collective)", share)
    # the API path: VolumetricModel.render_rays + backward() on each rank's share of the batches, ONE collective on the
    # peer-mapped buffer both .grad tensors are views of (PeerGradients), compared on every rank with the dense gradients of
    # an unsharded API render of the whole frame
    api = {}
    try:
        from thre3d_atom.modules.volumetric_model import VolumetricModel
        from thre3d_atom.rendering.volumetric.render_interface import Rays
        from thre3d_atom.thre3d_reprs.renderers import SHVoxGridRenderConfig, render_sh_voxel_grid
        from thre3d_atom.thre3d_reprs.voxels import VoxelGrid, VoxelSize
        from thre3d_atom.utils.imaging_utils import CameraBounds
        from voxe_b200.dist import PeerGradients

        grid = VoxelGrid(bench.dens.clone(), bench.feat.clone(), VoxelSize(*(w / n for w, n in zip(WL["world"], WL["dims"]))),
                         density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                         expected_density_scale=WL["density_scale"], tunable=True)
        vm = VolumetricModel(grid, render_sh_voxel_grid,
                             SHVoxGridRenderConfig(num_samples_per_ray=WL["S"], camera_bounds=CameraBounds(WL["near"], WL["far"]), white_bkgd=True,
                                                   perturb_sampled_points=False), device=device)
        o, d = bench.rays[pose]

        def api_grads(batches):
            for b0, b1 in batches:
                vm.render_rays(Rays(o[b0:b1], d[b0:b1])).colour.backward(bench.G[b0:b1])
            torch.cuda.synchronize(device)

        grid.densities.grad = grid.features.grad = None
        api_grads(bench.batches)
        want_d, want_f = grid.densities.grad.clone(), grid.features.grad.clone()
        if world > 1:
            grads = PeerGradients([grid.densities, grid.features])
            grads.zero()
            api_grads(bench.batches[rank::world])
            grads.allreduce()
            torch.cuda.synchronize(device)
            assert not grads.volume.failed(), "a peer did not arrive"
        errs = []
        for got, want in ((grid.densities.grad, want_d), (grid.features.grad, want_f)):
            t = torch.tensor([float((got - want).abs().max()) / float(want.abs().max())], device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            errs.append(float(t.item()))
        api = {"d_densities_max_abs_over_inf": errs[0], "d_features_max_abs_over_inf": errs[1], "ok": max(errs) <= 1e-5,
               "path": "render_rays + backward per batch on each rank's share, PeerGradients.allreduce(), against the unsharded API gradients"}
    except Exception as exc:  # noqa: BLE001
        api = {"error": str(exc)[:300]}
    results["API path (VolumetricModel + PeerGradients)"] = api
    # the API path with deferred gradients in a peer-mapped packed sink that keeps a brick trail: reduce_deferred() is the
    # brick-wise exchange, the optimiser-side hand-over follows the union of the flags (a part of the frame, two steps)
    if world > 1 and "error" not in api:
        sink = {}
        try:
            from voxe_b200.dist import VoxelGradAllReducer

            grid2 = VoxelGrid(bench.dens.clone(), bench.feat.clone(), VoxelSize(*(w / n for w, n in zip(WL["world"], WL["dims"]))),
                              density_preactivation=torch.nn.Identity(), density_postactivation=torch.nn.ReLU(),
                              expected_density_scale=WL["density_scale"], tunable=True)
            vm2 = VolumetricModel(grid2, render_sh_voxel_grid, vm.render_config, device=device)
            grid2.accumulate_render_gradients()
            spec = grid2.fused_spec()
            vol = PeerGradVolume(int(lib.voxe_packed_floats(spec.to_native())), device, multicast=world > 4)
            vol.adopt(grid2.render_gradient_accumulator, sparse_spec=spec)
            reducer = VoxelGradAllReducer([grid2.densities, grid2.features], grids=[grid2])
            errs = []
            for step, part in enumerate((bench.batches[: max(world, len(bench.batches) // 6)], bench.batches[len(bench.batches) // 2:][: 2 * world])):
                grid.densities.grad = grid.features.grad = None
                api_grads(part)  # unsharded, through the plain API path
                grid2.densities.grad = grid2.features.grad = None
                for b0, b1 in part[rank::world]:
                    vm2.render_rays(Rays(o[b0:b1], d[b0:b1])).colour.backward(bench.G[b0:b1])
                assert grid2.densities.grad is None
                reducer.reduce_deferred()
                grid2.materialize_render_gradients()
                torch.cuda.synchronize(device)
                assert not vol.failed(), "a peer did not arrive"
                for got, want in ((grid2.densities.grad, grid.densities.grad), (grid2.features.grad, grid.features.grad)):
                    t = torch.tensor([float((got - want).abs().max()) / float(want.abs().max()), float(vol.buffer.abs().max())], device=device)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    errs.append((float(t[0]), float(t[1])))
            sink = {"max_abs_over_inf": max(e for e, _ in errs), "sink_volume_left": max(v for _, v in errs),
                    "ok": max(e for e, _ in errs) <= 1e-5 and max(v for _, v in errs) == 0.0,
                    "path": "deferred gradients in a peer-mapped sink with a brick trail, VoxelGradAllReducer.reduce_deferred() = "
                            "voxe_allreduce_grads_peer_sparse, materialize = voxe_consume_grad along the union; two steps"}
        except Exception as exc:  # noqa: BLE001
            sink = {"error": str(exc)[:300]}
        results["API path (deferred sink + brick-wise exchange)"] = sink
    if rank == 0:
        ok = all(r.get("ok", True) and "error" not in r for r in results.values())
        print(json.dumps({"check": "rank-sum of render_bwd_kernel gradients", "n_gpus": world, "frame": f"pose {pose}, {len(bench.batches)} batches dealt round-robin",
                          "grad_inf_norm": scale, "collectives": results, "ok": ok}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def rank_cpu_slice(local, local_world):
    """This rank's share of the CPUs the process may run on, in units of physical cores (sysfs thread_siblings_list)."""
    allowed = sorted(os.sched_getaffinity(0))
    cores = {}
    for cpu in allowed:
        try:
            sib = Path(f"/sys/devices/system/cpu/cpu{cpu}/topology/thread_siblings_list").read_text().strip()
        except OSError:
            sib = str(cpu)
        cores.setdefault(sib, []).append(cpu)
    groups = sorted(cores.values(), key=lambda g: g[0])
    per = len(groups) // max(1, local_world)
    if per >= 1:
        return sorted(c for g in groups[local * per:(local + 1) * per] for c in g)
    per = len(allowed) // max(1, local_world)  # fewer cores than ranks: fall back to logical CPUs
    return allowed[local * per:(local + 1) * per] if per >= 1 else None


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU leg)")
    device = torch.device(f"cuda:{local}")
    torch.cuda.set_device(device)
    dist = None
    pinned_cpus = None
    if not args.no_pin and hasattr(os, "sched_setaffinity"):
        # one rank per GPU on a shared host: give every rank its own slice of the host's CPUs (what numactl / taskset
        # would do), so that eight Python loops, their autograd worker threads and NCCL helper threads do not migrate over
        # each other.  Slices are made of whole physical cores (both hyper-threads of a core go to the same rank): with
        # contiguous logical ids two ranks would share every core they own through its sibling thread.
        pinned_cpus = rank_cpu_slice(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        if pinned_cpus and args.cpus_per_rank > 0:
            pinned_cpus = pinned_cpus[:args.cpus_per_rank]
        if pinned_cpus:
            os.sched_setaffinity(0, pinned_cpus)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(WL["seed"] + rank)

    headline = args.workload == "cfg2"
    full = headline and not args.device_only  # --device-only: comparison runs keep the device leg, its parity check and roofline
    collective = args.collective
    peer_factory = None
    if world > 1 and collective != "nccl":
        try:
            peer_factory = make_peer_volume_factory(device, world, collective)
            probe = peer_factory(1024)  # fails here, on every rank alike, when the box cannot map peer memory
            collective = "voxe_allreduce_grads_peer (" + ("multimem.ld_reduce / multimem.st through the NVSwitch" if probe.multicast
                                                          else "peer loads / stores over NVLink") + ")"
            del probe
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] peer-mapped gradient volume unavailable ({str(exc)[:200]}); using ncclAllReduce", file=sys.stderr)
            peer_factory, collective = None, "nccl"
    if world > 1 and peer_factory is None:
        collective = "ncclAllReduce via torch.distributed"

    handover = args.handover or ("sparse" if args.workload == "cfg5" else "dense")
    if handover == "sparse" and world > 1 and peer_factory is None:
        handover = "dense"  # the brick-wise exchange is the library's own kernel; ncclAllReduce takes the whole volume
    bench = DeviceBench(device, rank, world, count_s_in=not (args.ncu or args.sweep), n_lanes=args.lanes,
                        kernel_jitter=args.jitter == "kernel", peer_volume=peer_factory, sparse=handover == "sparse")
    if bench.sparse and world > 1:
        collective = collective.replace("voxe_allreduce_grads_peer", "voxe_allreduce_grads_peer_sparse")
    barrier = (lambda: dist.barrier()) if world > 1 else None

    if args.ncu:  # profiler mode: eager launches of whole frames, nothing else (numbers printed here are NOT bench values)
        bench.frame_body(0, 0, "both")  # warm-up outside the profiled range
        torch.cuda.synchronize(device)
        torch.cuda.profiler.start()     # use with: ncu --profile-from-start off
        for k in range(args.steps):
            bench.frame_body((k + 1) % len(bench.poses), (k + 1) % bench.N_GRID_COPIES, "both")
            bench.unpack()
        torch.cuda.synchronize(device)
        torch.cuda.profiler.stop()
        print(json.dumps({"ncu_mode": True, "frames": args.steps, "kernels_per_frame": bench.kernels_per_step}))
        return

    def allreduce():
        if world > 1:
            if bench.peer_volume is not None and bench.sparse:
                bench.peer_volume.allreduce_sparse(bench.tag)  # flag union, then only the bricks some rank touched
            elif bench.peer_volume is not None:
                bench.peer_volume.allreduce()   # ONE all-reduce of the packed voxel gradients per step: the library's kernel
            else:
                dist.all_reduce(bench.packed_grad)

    def after_replay():
        allreduce()
        bench.unpack()

    if args.sweep:  # tuning mode: (L, rays per CTA, register cap) grid, isolated kernels + whole frame; not a bench line
        from voxe_b200 import _native as nat

        for cfg in args.sweep.split(";"):
            l, rpc, cap = (int(x) for x in cfg.split(","))
            try:
                nat.set_tuning(l, rpc, cap)
                bench.saved = torch.empty(int(bench.lib.voxe_saved_floats(bench.rd, WL["batch"])), device=device)
                bench.saved_lane = [bench.saved] + [torch.empty_like(bench.saved) for _ in range(bench.n_lanes - 1)]
                res = {}
                for what in ("fwd", "bwd", "both"):
                    gs = bench.capture(what, poses=sorted({0, 3 % len(bench.poses), 5 % len(bench.poses)}))
                    t_ms = bench.time_graphs(gs, 12, 3)
                    res[what] = round(1e3 * t_ms / (12 * len(bench.batches)), 2)
                    del gs
                    bench._keep = []
                print(json.dumps({"sweep": cfg, "us_per_batch": res}), flush=True)
            except Exception as exc:  # noqa: BLE001
                print(json.dumps({"sweep": cfg, "error": str(exc)[:200]}), flush=True)
        return

    random_views = bool(WL.get("random_views"))

    def timed(b, graphs, steps, warmup):
        """ms per step, max over ranks.  Weak-scaling workloads replay one graph (one frame) per step, rotating poses and
        packed-volume copies; cfg 4 (random_views) replays ALL of this rank's views per step before the one all-reduce."""
        if random_views:
            def step_all():
                for g in graphs:
                    g.replay()
                after_replay()
            for _ in range(warmup):
                step_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if barrier:
                barrier()
            torch.cuda.synchronize(device)
            e0.record()
            for _ in range(steps):
                step_all()
            e1.record()
            torch.cuda.synchronize(device)
            if barrier:
                barrier()
            ms = e0.elapsed_time(e1)
        else:
            ms = b.time_graphs(graphs, steps, warmup, after_replay=after_replay, barrier=barrier)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # strictly serialised variant first (one batch in flight), reported beside the headline
    serialized = None
    if bench.n_lanes > 1 and full:
        lanes = bench.n_lanes
        bench.n_lanes = 1
        g1 = bench.capture("both")
        ms1 = timed(bench, g1, args.steps, args.warmup)
        serialized = {"value": world * bench.R / (ms1 * 1e-3), "ms_per_step": ms1,
                      "note": "one batch in flight: fwd(k) -> bwd(k) -> fwd(k+1) ... on a single stream"}
        del g1
        bench.n_lanes = lanes

    graphs = bench.capture("both")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_per_step = timed(bench, graphs, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    rays_per_step = (bench.total_views * bench.R) if random_views else world * bench.R
    rays_per_s = rays_per_step / (ms_per_step * 1e-3)
    last = ((args.warmup + args.steps - 1) * world + rank) % len(graphs)  # graph index = pose of the last timed frame
    parity = None
    if rank == 0 and (headline or args.parity) and not random_views:
        parity = parity_of_timed_leg(bench, last, last % bench.N_GRID_COPIES, world)
    if world > 1:
        dist.barrier()

    # where a step of the sparse hand-over spends its time: the stages of the last frames again, one at a time (every rank
    # takes part in the exchange; rank 0 reports)
    breakdown = None
    if bench.sparse:
        stages = {"zero_dense_grads_and_flags": [], "fwd_bwd": [], "allreduce_sparse": [], "consume": []}
        for rep in range(3):
            pose = (last + rep) % len(bench.poses)
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            if barrier:
                barrier()
            torch.cuda.synchronize(device)
            marks[0].record()
            bench.d_dens.zero_()
            bench.d_feat.zero_()
            bench.touched.zero_()
            marks[1].record()
            bench.frame_body(pose, 0, "both", zero=False)
            marks[2].record()
            allreduce()
            marks[3].record()
            bench.unpack()
            marks[4].record()
            torch.cuda.synchronize(device)
            for name, a, b in zip(stages, marks[:-1], marks[1:]):
                stages[name].append(a.elapsed_time(b))
        breakdown = {k: round(1e3 * min(v), 1) for k, v in stages.items()}
        breakdown["unit"] = "us, best of 3 eager passes on rank 0 (allreduce_sparse includes waiting for the slowest rank's backward)"

    # the same frame on a Softplus field (the reference scripts' default): every in-grid sample scatters
    softplus = None
    if full:
        twin = DeviceBench(device, rank, world, count_s_in=False, n_lanes=args.lanes, kernel_jitter=args.jitter == "kernel",
                           postact="softplus", share=bench)
        tg = twin.capture("both")
        tms = timed(twin, tg, max(10, args.steps // 4), 3)
        softplus = {"value": world * bench.R / (tms * 1e-3), "ms_per_step": tms, "postact": "softplus"}
        if rank == 0:
            last_t = ((3 + max(10, args.steps // 4) - 1) * world + rank) % len(tg)
            softplus["parity"] = parity_of_timed_leg(twin, last_t, last_t % bench.N_GRID_COPIES, world)
        if world > 1:
            dist.barrier()

    # dominant kernel alone (rank 0; other ranks idle-wait at the barrier below)
    peak, peak_note = measured_hbm_peak()
    roof = None
    if rank == 0:
        s_in_mean = sum(t for t, _ in bench.s_in) / len(bench.s_in)
        n_b = len(bench.batches)

        def kernel_us(b, what):
            gs = b.capture(what, poses=sorted({0, 3 % len(b.poses), 5 % len(b.poses)}))
            b.world, saved = 1, b.world
            t_ms = b.time_graphs(gs, max(8, args.steps), 3)
            b.world = saved
            return 1e3 * t_ms / (max(8, args.steps) * n_b)

        def moved(b, what, us):
            """Bytes one launch of the kernel moves (frame average), per the kernel's own counters."""
            n_in, n_scatter = b.scatter_stats(pose=0)
            s_in_pose0 = b.s_in[0][0] if b.s_in else bench.s_in[0][0]
            if what == "fwd":  # every in-grid sample: 8-corner gather + one saved vector
                per_frame = s_in_mean * (MOVED_GATHER + MOVED_SAVED_VECTOR) + bench.R * moved_bytes_per_ray("fwd")
                frac_scatter = None
            else:              # processed samples reload one vector; only samples with a non-zero gradient scatter
                frac_processed, frac_scatter = n_in / s_in_pose0, n_scatter / s_in_pose0
                per_frame = s_in_mean * (frac_processed * MOVED_SAVED_VECTOR + frac_scatter * MOVED_GATHER) + bench.R * moved_bytes_per_ray("bwd")
            return per_frame / n_b, frac_scatter

        bwd_us, fwd_us = kernel_us(bench, "bwd"), kernel_us(bench, "fwd")
        bwd_bytes, frac_scatter = moved(bench, "bwd", bwd_us)
        fwd_bytes, _ = moved(bench, "fwd", fwd_us)
        model_bwd = (s_in_mean * MODEL_BYTES_PER_SAMPLE_BWD + bench.R * MODEL_BYTES_PER_RAY) / n_b
        model_fwd = (s_in_mean * MODEL_BYTES_PER_SAMPLE_FWD + bench.R * MODEL_BYTES_PER_RAY) / n_b
        model_step = s_in_mean * (MODEL_BYTES_PER_SAMPLE_FWD + MODEL_BYTES_PER_SAMPLE_BWD) + bench.R * 96
        moved_step = (bwd_bytes + fwd_bytes) * n_b
        gbs = lambda nbytes, us: nbytes / (us * 1e-6) / 1e9  # noqa: E731
        l2_peak, l2_mb = l2_probe(device)
        # the committed captures: cfg 2 (<0, 3, 96>) and cfg 5 (<2, 3, 128>) kernels
        has_capture = args.workload in ("cfg2", "cfg5") and args.batch == 0
        cap, cap_src = ncu_capture(f"render_bwd_kernel<{WL['sh_degree']}, 3") if has_capture else (None, None)
        cap_f, _ = ncu_capture(f"render_fwd_kernel<{WL['sh_degree']}, 3") if has_capture else (None, None)
        step_us = ms_per_step * 1e3 / (len(graphs) if random_views else 1)
        roof = {
            "bound": "hbm", "achieved": round(gbs(bwd_bytes, bwd_us), 1), "peak": peak, "unit": "GB/s",
            "frac": round(gbs(bwd_bytes, bwd_us) / peak, 4), "peak_source": peak_note,
            "kernel": f"render_bwd_kernel<DEG={WL['sh_degree']},NCOL=3> ({bench.postact})", "us_per_launch": round(bwd_us, 2),
            "algorithmic_bytes_per_launch": round(bwd_bytes),
            "bytes_billed": "per in-grid sample the kernel processed: 16 B saved-vector reload + (fraction that scatters) x 8 corners x "
                            f"{MOVED_GATHER // 8} B RED payload; per ray: rays + dL/dcolour + segment summaries; fractions counted by the kernel "
                            "(VoxeRenderDesc.stats)",
            "scatter_fraction": None if frac_scatter is None else round(frac_scatter, 4),
            "intra_warp_duplicates": getattr(bench, "merge_stats", None),
            # the other ceiling of this kernel: an SM issues vector REDs at ~1.29 clk per lane (B300_MICROARCH.md "Atomics",
            # spread addresses); the floor below is lane-REDs per launch x 1.29 clk / (SMs x SM clock)
            "red_issue": red_issue_floor(getattr(bench, "merge_stats", None), n_b, bwd_us, clocks),
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture (the grid is
            # L2-resident at 160^3, hence far below the algorithmic bytes)
            "traffic": None if not cap else cap.get("dram_bytes"),
            "traffic_source": cap_src,
            "model_frac": round(gbs(model_bwd, bwd_us) / peak, 4),
            "model_note": "SURVEY.md 8d contract (backward billed 2 x 8 corners x (F+1) x 4 B per in-AABB sample: a re-gather the kernel "
                          "replaced by the 16-byte reload, and a scatter for every sample) -- kept for continuity, not a ceiling",
            "l2": {"peak_measured": round(l2_peak, 1), "unit": "GB/s", "how": f"torch.mul between two {l2_mb} MB L2-resident buffers (best of 8..40 MB), read+write bytes / time",
                   "lts_bytes_per_launch": None if not cap else cap.get("lts_bytes"),
                   "achieved": None if not cap or not cap.get("lts_bytes") else round(gbs(cap["lts_bytes"], bwd_us), 1),
                   "frac": None if not cap or not cap.get("lts_bytes") else round(gbs(cap["lts_bytes"], bwd_us) / l2_peak, 4)},
            "fwd_kernel": {"us_per_launch": round(fwd_us, 2), "algorithmic_bytes_per_launch": round(fwd_bytes),
                           "achieved": round(gbs(fwd_bytes, fwd_us), 1), "frac": round(gbs(fwd_bytes, fwd_us) / peak, 4),
                           "model_frac": round(gbs(model_fwd, fwd_us) / peak, 4),
                           "traffic": None if not cap_f else cap_f.get("dram_bytes"),
                           "l2_frac": None if not cap_f or not cap_f.get("lts_bytes") else round(gbs(cap_f["lts_bytes"], fwd_us) / l2_peak, 4)},
            "step": {"achieved": round(gbs(moved_step, step_us), 1), "frac": round(gbs(moved_step, step_us) / peak, 4),  # per GPU
                     "model_frac": round(gbs(model_step, step_us) / peak, 4),
                     "bytes_per_ray": round(moved_step / bench.R, 1), "s_in_per_ray": round(s_in_mean / bench.R, 2)},
        }
        if softplus is not None:
            t_bwd, t_fwd = kernel_us(twin, "bwd"), kernel_us(twin, "fwd")
            t_bytes, t_frac = moved(twin, "bwd", t_bwd)
            softplus["bwd_kernel"] = {"us_per_launch": round(t_bwd, 2), "scatter_fraction": round(t_frac, 4), "algorithmic_bytes_per_launch": round(t_bytes),
                                      "achieved": round(gbs(t_bytes, t_bwd), 1), "frac": round(gbs(t_bytes, t_bwd) / peak, 4)}
            softplus["fwd_kernel"] = {"us_per_launch": round(t_fwd, 2)}
            softplus["bwd_kernel"]["intra_warp_duplicates"] = getattr(twin, "merge_stats", None)
            softplus["bwd_kernel"]["red_issue"] = red_issue_floor(getattr(twin, "merge_stats", None), n_b, t_bwd, clocks)
    if world > 1:
        dist.barrier()

    e2e = None
    if full:
        n_e2e, w_e2e = max(2, min(args.steps, args.e2e_steps)), min(args.warmup, 3)
        e2e = e2e_leg(device, rank, world, n_e2e, w_e2e, dist, peer_volume=peer_factory)
        variants = {
            "cuda_graph": (dict(graph=True, peer_volume=peer_factory), "the same API calls of a frame captured once with torch.cuda.graph and replayed (stock torch; "
                                             "H2D of the frame's inputs and D2H of its results stay in the timed region)"),
            "calling_thread_engine": (dict(engine_threads=False, peer_volume=peer_factory), "same loop under torch.autograd.set_multithreading_enabled(False): backward() "
                                                                  "runs on the calling thread (a caller-side switch)"),
            "frame_backward": (dict(frame_backward=True, peer_volume=peer_factory),
                               "40 x render_rays(4096 rays), then ONE torch.autograd.backward over the 40 outputs (the frame's loss is the sum over "
                               "its batches): every batch is still rendered and differentiated, the autograd engine is entered once per frame"),
            "deferred_grads": (dict(deferred=True, peer_volume=peer_factory), "same loop with VoxelGrid.accumulate_render_gradients(): gradients "
                                                                              "materialised once per frame"),
            "whole_frame_call": (dict(whole_frame=True, peer_volume=peer_factory), "not the headline workload: the same frame as ONE render_rays(160000 rays) + ONE backward() "
                                                         "(the SDS edit loop's calling pattern)"),
        }
        for name, (kw, note) in variants.items():
            try:
                r = e2e_leg(device, rank, world, n_e2e, w_e2e, dist, **kw)
                e2e[name] = {"value": r["value"], "ms_per_step": r["ms_per_step"], "note": note}
            except Exception as exc:  # noqa: BLE001 -- a variant that cannot run is reported, the headline stands
                e2e[name] = {"error": str(exc)[:300], "note": note}
        e2e["host_affinity"] = ("unpinned (the process may run on every CPU of the host)" if not pinned_cpus
                                else f"every rank pinned to {len(pinned_cpus)} CPU(s) of its own slice of the host (os.sched_setaffinity)")
        e2e["collective"] = None if world == 1 else (
            "one per frame: voxe_allreduce_grads_peer in place on the peer-mapped buffer both .grad tensors are views of (voxe_b200.dist.PeerGradients)"
            if peer_factory is not None else "one per frame: VoxelGradAllReducer (flat staging buffer of both dense gradients, ncclAllReduce)")

    cpu = gpu_baseline = None
    if rank == 0 and world == 1 and full:
        if not args.no_cpu:
            r = reference_subprocess("cpu", 8, 1)
            if "error" in r:
                cpu = {"error": r["error"]}
            else:
                cpu = dict(r["cpu_baseline"])
        r = reference_subprocess("cuda", 10, 3)
        if "error" in r:
            gpu_baseline = {"error": r["error"]}
        else:
            gpu_baseline = {"value": r["value"], "unit": "rays/s", "ms_per_batch": r["ms_per_step"], "kind": r["cpu_baseline"]["kind"],
                            "what": "the same 4096-ray batches fwd+bwd through stock ATen ops on this GPU (BASELINE.md: the bar the kernels have to beat)",
                            "sample": r["cpu_baseline"]["sample"],
                            "ours_serialized_over_baseline": None if not serialized else round(serialized["value"] / r["value"], 1)}

    fused_step = None
    if rank == 0 and world == 1 and full:
        fused_step = fused_step_leg(device, peak)

    inference = None
    if rank == 0 and world == 1 and full:
        inference = inference_leg(device)

    regularizers = sampler = None
    if rank == 0 and world == 1 and full:
        regularizers = regularizers_leg(device, peak)
        sampler = sampler_leg(device)

    if rank == 0:
        if random_views:
            step = (f"{bench.total_views} get_random_pose views of {WL['height']}x{WL['width']} dealt round-robin over {world} rank(s), each view one "
                    f"launch pair (fwd, bwd) accumulating into the packed gradient volume, then ONE all-reduce ({bench.packed_grad.numel() * 4 / 1e6:.0f} MB) "
                    "+ unpack; total work fixed as N grows")
        else:
            step = (f"one {WL['height']}x{WL['width']} frame = {len(bench.batches)} batches x (fwd, bwd; jitter "
                    f"{'generated in-kernel' if bench.kernel_jitter else 'drawn by torch.rand'})"
                    + (" + zero-fill of the dense gradients + brick-flag hand-over (voxe_consume_grad)" if bench.sparse else " + grad zero-fill + unpack")
                    + ((f" + brick-wise all-reduce of the packed voxel grads (flag union, then the touched bricks of {bench.packed_grad.numel() * 4 / 1e6:.0f} MB)"
                        if bench.sparse else f" + 1 all-reduce of the packed voxel grads ({bench.packed_grad.numel() * 4 / 1e6:.0f} MB)") if world > 1 else ""))
        line = {
            "metric": "rays/s fwd+bwd, 160^3 SH-0 grid, 400x400 render", "value": rays_per_s, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if random_views else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WL["name"], "step": step,
                       "batches_in_flight": f"{bench.n_lanes} (each launch is one <={WL['batch']}-ray batch with its own workspace; gradients accumulate "
                                            "over the frame, so batch k+1's forward does not wait for batch k's backward)",
                       "l2": f"{bench.N_GRID_COPIES} rotating packed-volume copies ({bench.N_GRID_COPIES * bench.packed[0].numel() * 4 / 1e6:.0f} MB) + "
                             f"{bench.packed_grad.numel() * 4 / 1e6:.0f} MB gradient volume + per-batch workspaces > 126 MB L2; {len(bench.poses)} poses rotate",
                       "parallelism": f"ray/view data parallel x{world}", "timing": "CUDA events around K graph replays, max over ranks",
                       "collective": None if world == 1 else collective,
                       "gradient_handover": "sparse (brick flags)" if bench.sparse else "dense",
                       "touched_brick_fraction": getattr(bench, "touched_fraction", None),
                       "step_breakdown": breakdown,
                       "host_cpus_per_rank": len(pinned_cpus) if pinned_cpus else "unpinned"},
            "e2e": e2e, "gpu_launches": args.steps * ((bench.kernels_per_step - 1) * (len(graphs) if random_views else 1) + 1
                                                         + (1 if bench.peer_volume is not None else 0)), "roofline": roof,
            "clocks": clocks, "parity": parity,
        }
        if softplus:
            line["value_softplus"] = softplus["value"]
            line["softplus"] = softplus
        if cpu:
            line["cpu_baseline"] = cpu
        if gpu_baseline:
            line["gpu_baseline"] = gpu_baseline
        if fused_step:
            line["fused_step"] = fused_step
        if serialized:
            line["serialized"] = serialized
        if inference:
            line["inference"] = inference
        if regularizers:
            line["regularizers"] = regularizers
        if sampler:
            line["batch_sampler"] = sampler
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--ref-device", choices=["cpu", "cuda"], default="cpu",
                    help="--impl reference: cpu = the reference arm (host cores); cuda = the stock-ATen GPU baseline")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--device-only", action="store_true", help="comparison runs: skip the e2e / baseline / side legs (not the driver's line)")
    ap.add_argument("--ncu", action="store_true", help="profiler mode: run --steps eager frames and exit")
    ap.add_argument("--check", action="store_true", help="rank-sum check of the CUDA path's gradients through every collective; not a bench line")
    ap.add_argument("--hybrid", action="store_true", help="--check: also time launches split between the multicast and the plain peer path")
    ap.add_argument("--parity", action="store_true", help="run the in-run oracle check for workloads other than cfg2 as well")
    ap.add_argument("--workload", choices=["cfg2", "cfg3", "cfg4", "cfg5"], default="cfg2",
                    help="cfg2 is the headline (the line the driver reads); the others are recorded in DESIGN.md / profiles/")
    ap.add_argument("--collective", choices=["peer", "peer-p2p", "nccl"], default="peer",
                    help="N > 1: voxe_allreduce_grads_peer on a peer-mapped gradient volume (multimem when the box has NVLS; peer-p2p "
                         "forces plain peer loads/stores), or ncclAllReduce through torch.distributed")
    ap.add_argument("--handover", choices=["dense", "sparse"], default="",
                    help="gradient hand-over of the device leg: 'dense' = zero-fill + whole-volume all-reduce + unpack; 'sparse' = follow the "
                         "backward's brick flags (default for cfg5, whose batch touches a few percent of a 15 GB volume)")
    ap.add_argument("--lanes", type=int, default=4, help="ray batches in flight within a frame (streams)")
    ap.add_argument("--jitter", choices=["kernel", "buffer"], default="kernel",
                    help="stratified jitter of the device leg: generated inside the kernels (counter-based hash) or torch-drawn [R,S] buffers")
    ap.add_argument("--sweep", type=str, default="", help="tuning sweep: 'L,rpc,cap;L,rpc,cap;...'")
    ap.add_argument("--tune", type=str, default="", help="L,rpc,regcap launch-shape override for tuning runs")
    ap.add_argument("--cpus-per-rank", type=int, default=1,
                    help="N > 1: CPUs of its slice a rank is pinned to (0 = the whole slice).  Default 1: on these (virtualised) hosts the "
                         "hand-off between the Python thread and torch's autograd worker thread costs ~8 us per hop across cores and "
                         "degrades when eight ranks do it at once (5.1 ms per frame at N = 8 on 4-CPU slices, 3.1 ms on one CPU)")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not pin each rank to its own slice of the host CPUs")
    ap.add_argument("--shuffle-rays", action="store_true",
                    help="measurement variant: every frame's rays in random order (incoherent batches, as random training batches are)")
    ap.add_argument("--batch", type=int, default=0, help="rays per launch override (tuning / ray-batch sweeps; not a bench line)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.batch > 0:
        WL["batch"] = args.batch
    if args.shuffle_rays:
        WL["shuffle_rays"] = True
        WL["name"] += " [rays of every frame shuffled]"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.tune:
        from voxe_b200 import _native as nat

        nat.set_tuning(*(int(x) for x in args.tune.split(",")))
    if args.impl == "reference":
        run_reference(args)
    elif args.check:
        run_check(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
