#!/usr/bin/env python
"""bench.py — RGB-D frames/s (and Mvoxels-updated/s, % of HBM roofline) of the compute() hot path.

Workload: synthetic orbiting-camera RGB-D stream (scene S2 of SURVEY.md §8d, replica.cfg parameters),
640x480 / 1000 frames per orbit at N = 1 (BASELINE.json configs[1]), 1280x960 / 2000 frames per orbit
at N > 1 (configs[3]) unless --width/--height say otherwise. One "step" = one frame through compute():
block allocation -> visibility -> TSDF fusion + garbage collection (one persistent kernel, k_frame).

  value  : frames/s with the frames already resident in HBM (device pointers handed to the C ABI),
           L2 flushed before every step, per-step CUDA-event time summed, max over ranks.
  e2e    : frames/s through the public GeoWrapper API with HOST buffers, pipelined the way a streaming
           caller uses it: per step setCurrPose + setDepthImage + setRGBImage (H2D from page-locked
           frames, mrh_set_ingest_mode 2) + compute() + the counters of the previous frame read back
           (mrh_get_stats_pipelined); the transfer of frame k+1 overlaps the kernel of frame k.
  roofline: k_frame, algorithmic bytes per frame / CUDA-event kernel time vs MEASURED_PEAKS.json.
  cpu_baseline: the CPU restatement (oracle/) with OpenMP on the host cores, bounded sample.
  lidar, mesh_50M (N = 1): BASELINE configs[2] and configs[4] through the same public API (tools/bench_lidar.py,
           tools/bench_mesh.py), each next to the reference kernels' time on the same input.

`--impl reference` runs the UNMODIFIED reference kernels (oracle/_ref/libref_harness.so, compiled
from /root/reference for sm_100a) on the same stream, host buffers in, the way GeoWrapper::compute
drives them: `value` is the CUDA-event time of its integrate() section (device time, like ours),
`e2e` its wall-clock frames/s with host buffers. N>1 (torchrun): the workload is BASELINE configs[3]
(1280x960, 2000-frame orbit - NOT the 640x480 stream of `--gpus 1`); the map is sharded by hash-bucket
range, every rank uploads its band of rows of each frame and an NCCL all-gather completes it
(sharding.scatter_ingest_frame); every rank allocates / fuses only the blocks it owns (starve frames
min-reduce the z-buffer over the ranks). More keys at N>1: `replica_streams` (one unsharded stream per
GPU, no collective), `single_gpu_same_workload` (what one GPU does on this line's stream: the
denominator for a scaling figure on this workload) and `sharded_mesh` (boundary exchange + marching
cubes in place + soup gather + weld, timed). stdout carries exactly one JSON line; everything else
goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_SDF_BLOCKS = 500000  # the reference's own test sizing (tests/test_hash_utils.cu:175-190)
HASH_NUM_BUCKETS = 250000
L2_FLUSH_BYTES = 256 << 20
ALL_CPUS = os.sched_getaffinity(0)  # before bind_near_gpu narrows it
STATS_LAG = int(os.environ.get("MRH_BENCH_STATS_LAG", "2"))  # e2e passes read every frame's counters this many frames after submitting it
COUNTERS_BYTES = 144  # the part of mrh::Counters that getStats() / mrh_get_stats_pipelined read back


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=0, help="default: 640 at --gpus 1, 1280 at --gpus > 1 (BASELINE configs[1] / configs[3])")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--orbit-frames", type=int, default=0, help="frames per full orbit of scene S2 (default 1000 / 2000)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the lidar / mesh_50M keys (N = 1)")
    args = ap.parse_args()
    multi = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) > 1
    if not args.width:
        args.width = 1280 if multi else 640
    if not args.height:
        args.height = args.width * 3 // 4
    if not args.orbit_frames:
        args.orbit_frames = 2000 if args.width >= 1280 else 1000
    return args


def workload(args):
    """The same string in both arms (the driver compares config.workload)."""
    return f"S2 orbiting-camera RGB-D stream {args.width}x{args.height} ({args.orbit_frames} frames/orbit), replica.cfg parameters (BASELINE configs[{3 if args.width >= 1280 else 1}])"


class ClockSampler:
    """SM clock and throttle reasons sampled during the timed regions (B200_PROFILING.md), through
    NVML in a background thread (nvidia-smi polling takes driver locks that stall concurrent CUDA API
    calls: at 100 ms it more than doubled the per-frame wall time of the call-heavy reference arm)."""

    def __init__(self, gpu_index, period_s=0.02):
        self.idx = gpu_index
        self.period = period_s
        self.rows = []  # (time, sm_mhz, sm_max_mhz, [reasons])
        self._stop = threading.Event()
        self._thread = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.idx]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.idx
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                "hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4),
            }
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(get_reasons(h))
                        self.rows.append((time.time(), sm, mx, [n for n, bit in names.items() if mask & bit]))
                    except Exception:
                        pass
                    self._stop.wait(self.period)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._thread = None

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        for ts, s, m, rs in self.rows:
            if not any(a - self.period <= ts <= b + self.period for a, b in windows):
                continue
            sm.append(s)
            mx = max(mx, m)
            reasons.update(rs)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def make_stream(args, n, device):
    """Frames 0..n-1 of scene S2 rendered on `device` (torch), plus poses."""
    import torch

    from mrhash_b200 import synth

    depth = torch.empty((n, args.height, args.width), dtype=torch.float32, device=device)
    rgb = torch.empty((n, args.height, args.width, 3), dtype=torch.uint8, device=device)
    poses = []
    for k in range(n):
        t, q, R = synth.orbit_pose(k, args.orbit_frames)
        d, c = synth.render_rgbd_torch(R, t, args.width, args.height, device=device)
        depth[k], rgb[k] = d, c
        poses.append((t, q))
    return depth, rgb, poses


def new_map(args, rank, world, device_index, max_num_triangles=1):
    from mrhash_b200 import GeoWrapper, synth

    p = dict(synth.REPLICA_PARAMS)
    g = GeoWrapper(**p, num_sdf_blocks=NUM_SDF_BLOCKS, hash_num_buckets=HASH_NUM_BUCKETS, max_num_triangles=max_num_triangles, device=device_index, shard_rank=rank, shard_world=world)
    fx, fy, cx, cy = synth.intrinsics(args.width, args.height)
    g.setCamera(fx, fy, cx, cy, args.height, args.width, p["min_depth"], p["max_depth"], 0)
    return g


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(args, depth_h, rgb_h, poses, budget_s):
    """The CPU restatement with all host threads on the first frames of the same stream."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle

    from mrhash_b200 import synth

    bound = os.sched_getaffinity(0)
    os.sched_setaffinity(0, ALL_CPUS)  # the CPU arm gets every core of the box, not only the GPU's NUMA node
    cores = len(os.sched_getaffinity(0))
    p = dict(synth.REPLICA_PARAMS)
    o = Oracle(p, 100000, 50000, threads=cores)
    fx, fy, cx, cy = synth.intrinsics(args.width, args.height)
    o.set_camera(fx, fy, cx, cy, args.height, args.width, p["min_depth"], p["max_depth"], 0)
    t0 = time.perf_counter()
    n = 0
    while n < len(poses) and (time.perf_counter() - t0 < budget_s or n < 3):
        t, q = poses[n]
        o.compute_rgbd(synth.quat_to_matrix_f32(t, q), depth_h[n], rgb_h[n])
        n += 1
    dt = time.perf_counter() - t0
    os.sched_setaffinity(0, bound)
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port", "sample": f"first {n} frames of the same {args.width}x{args.height} stream ({dt:.1f} s), oracle/mrh_oracle.c with OpenMP over pixel rows / blocks"}


def run_reference(args, rank, world, numa=""):
    """The reference's own kernels (oracle/_ref), driven like GeoWrapper::compute, host buffers in."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from oracle_lib import RefCuda, ref_available

    from mrhash_b200 import synth

    base = {"impl": "reference", "metric": "rgbd_frames_per_sec", "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True}
    if not ref_available() or not torch.cuda.is_available():
        print(json.dumps({**base, "unavailable": "oracle/_ref/libref_harness.so missing or no CUDA device"}), flush=True)
        return
    n = args.warmup + args.steps
    depth, rgb, poses = make_stream(args, n, "cuda:0")
    depth_h, rgb_h = depth.cpu().numpy(), rgb.cpu().numpy()
    del depth, rgb
    p = dict(synth.REPLICA_PARAMS)
    fx, fy, cx, cy = synth.intrinsics(args.width, args.height)
    # silence the reference's per-frame prints (stdout of this process must stay one JSON line)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1), os.dup(2)
    os.dup2(devnull, 1), os.dup2(devnull, 2)
    try:
        cwd = os.getcwd()
        os.makedirs("/tmp/mrh_ref_run", exist_ok=True)
        os.chdir("/tmp/mrh_ref_run")  # the reference writes ./<name>.txt profiler logs
        r = RefCuda(p, NUM_SDF_BLOCKS, HASH_NUM_BUCKETS)
        r.set_camera(fx, fy, cx, cy, args.height, args.width, p["min_depth"], p["max_depth"], 0)
        sampler = ClockSampler(0)
        sampler.start()
        for k in range(args.warmup):
            r.compute_rgbd(synth.quat_to_matrix_f32(*poses[k]), depth_h[k], rgb_h[k])
        torch.cuda.synchronize()
        w0 = time.time()
        t0 = time.perf_counter()
        integ_ms = 0.0
        for k in range(args.warmup, n):
            r.compute_rgbd(synth.quat_to_matrix_f32(*poses[k]), depth_h[k], rgb_h[k])
            integ_ms += r.last_integrate_ms()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        w1 = time.time()
        sampler.stop()
        occupied = r.occupied()
        os.chdir(cwd)
    finally:
        os.dup2(saved[0], 1), os.dup2(saved[1], 2)
    fps_wall = args.steps / dt
    fps_dev = args.steps / (integ_ms * 1e-3) if integ_ms > 0 else None
    line = {
        **base,
        "value": fps_dev if fps_dev else fps_wall,
        "value_kind": "device time: CUDA events around VoxelContainer::integrate (voxel_data_structures.cpp:94-109), inputs already uploaded" if fps_dev else "wall clock",
        "ms_per_step": (integ_ms / args.steps) if fps_dev else 1e3 * dt / args.steps,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload(args), "num_sdf_blocks": NUM_SDF_BLOCKS, "hash_num_buckets": HASH_NUM_BUCKETS, "l2": "not flushed (host-driven, synchronous reference)", "host_placement": numa, "parallelism": "1 GPU (the reference has no multi-GPU path)"},
        "integrate_only_fps": fps_dev,
        "wall_clock_fps": fps_wall,
        "visible_blocks_last_frame": occupied,
        "cpu_baseline": {"value": fps_wall, "unit": "frames/s", "kind": "reference", "cores": 1, "sample": f"{args.steps} frames; the reference has no CPU path: this is its own CUDA code (oracle/_ref, -arch=sm_100a) driven by one host thread"},
        "e2e": {"value": fps_wall, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "what": "wall clock, host buffers in, as GeoWrapper::compute drives the kernels"},
        "clocks": sampler.summary([(w0, w1)]),
        "gpu_launches": None,
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(gpu_index):
    """Run this process on the CPUs of the GPU's NUMA node, so that the page-locked frames it allocates (first
    touch) sit behind the PCIe root the GPU hangs on: the same 1.2 MB upload was measured at 47 GB/s from the
    near node and 20-31 GB/s from the far one. Returns what was done, for the JSON line. MRH_BENCH_NUMA=0 skips it."""
    if os.environ.get("MRH_BENCH_NUMA", "1") == "0":
        return "not bound (MRH_BENCH_NUMA=0)"
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "not bound (the device reports no NUMA node)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        done = []
        # memory first: prefer the node for every later allocation of this process (MPOL_PREFERRED), whether or
        # not its CPUs are ours to run on
        try:
            import ctypes

            mask = ctypes.c_ulong(1 << node)
            if ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), 8 * ctypes.sizeof(mask)) == 0:  # set_mempolicy
                done.append(f"memory preferred on NUMA node {node}")
        except Exception:
            pass
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            done.append(f"{len(cpus)} CPUs of NUMA node {node}")
        return (", ".join(done) + f" (GPU {gpu_index})") if done else f"not bound (node {node}: no usable CPU, no memory policy)"
    except Exception as exc:  # no NVML / no sysfs: run wherever the scheduler put us
        return f"not bound ({type(exc).__name__})"


def main():
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (the library's
    # reference-style progress prints, NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_near_gpu(local)  # both arms, before anything allocates host memory
    print(f"[bench] host placement: {numa}", file=sys.stderr, flush=True)
    if args.impl == "reference":
        run_reference(args, rank, world, numa)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (mrhash_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return x

    W, K = args.warmup, args.steps
    n = W + K
    depth, rgb, poses = make_stream(args, n, dev)
    P = args.width * args.height
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    def prof_range(on):
        # MRH_PROFILE_RANGE=1 + `ncu --profile-from-start off`: the launch list then holds the timed
        # regions only (the synthetic stream is rendered by ~1500 torch launches before them)
        if os.environ.get("MRH_PROFILE_RANGE"):
            rt = torch.cuda.cudart()
            rt.cudaProfilerStart() if on else rt.cudaProfilerStop()

    # ---------------- pass A: device-resident inputs, L2 flushed before every step -----------------
    g = new_map(args, rank, world, local)
    stream = torch.cuda.ExternalStream(g.cudaStream(), device=dev)

    from mrhash_b200 import sharding

    def run_frame(geo):
        # one rank of a sharded map: compute() plus, on starve frames, the z-buffer min-reduction
        if world > 1 and geo.shardWorld() > 1:
            sharding.compute_sharded(geo)
        else:
            geo.compute()

    def step_device(k):
        g.setCurrPose(*poses[k])
        g.setDepthImageDevice(depth[k].data_ptr(), args.height, args.width)
        g.setRGBImageDevice(rgb[k].data_ptr(), args.height, args.width)
        run_frame(g)

    for k in range(W):
        step_device(k)
    g.synchronize()
    if world > 1:
        # The only collective on the integration path is the z-buffer min-reduction of a starve frame
        # (every n_frames_invalidate_voxels-th frame: none falls into the warm-up). NCCL sets up the
        # channels of a (collective, size) pair on first use - tens of milliseconds at 8 ranks - so that
        # first use happens here, not inside a timed step.
        zb = torch.full((P,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
        for _ in range(2):
            dist.all_reduce(zb, op=dist.ReduceOp.MIN)
        del zb
        barrier()
    g.resetStats()
    launches0 = g.launchCount()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    windows.append([time.time(), None])
    prof_range(True)
    with torch.cuda.stream(stream):
        for i in range(K):
            flush.zero_()
            ev[i][0].record(stream)
            step_device(W + i)
            ev[i][1].record(stream)
    g.synchronize()
    barrier()
    prof_range(False)
    windows[-1][1] = time.time()
    ms_flushed = sum(a.elapsed_time(b) for a, b in ev)
    ms_flushed = max_over_ranks(ms_flushed)
    st = g.getStats()
    launches = g.launchCount() - launches0
    v_upd = sum_over_ranks(st["voxels_updated"])
    b_vis = sum_over_ranks(st["blocks_visible"])
    b_new = sum_over_ranks(st["blocks_new"])
    live = sum_over_ranks(st["live_blocks"])
    assert st["dropped_heap"] == 0 and st["dropped_table"] == 0, st
    g.close()

    # ---------------- pass A': same, L2 as the stream leaves it (whole region, one event pair) -------
    g = new_map(args, rank, world, local)
    stream = torch.cuda.ExternalStream(g.cudaStream(), device=dev)
    for k in range(W):
        step_device(k)
    g.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    windows.append([time.time(), None])
    prof_range(True)
    e0.record(stream)
    for i in range(K):
        step_device(W + i)
    e1.record(stream)
    g.synchronize()
    barrier()
    prof_range(False)
    windows[-1][1] = time.time()
    ms_stream = max_over_ranks(e0.elapsed_time(e1))
    g.close()

    # ---------------- pass B: end to end through the public API, host buffers ----------------------
    # N > 1 (8.6 MB frames, one copy of the stream per rank): the end-to-end pass runs at most 400 timed
    # steps, so that eight ranks do not page-lock 70 GB of host memory between them
    Ke = K if world == 1 else min(K, 400)
    ne = W + Ke
    depth_h = torch.empty((ne,) + tuple(depth.shape[1:]), dtype=depth.dtype, pin_memory=True)
    rgb_h = torch.empty((ne,) + tuple(rgb.shape[1:]), dtype=rgb.dtype, pin_memory=True)
    depth_h.copy_(depth[:ne])
    rgb_h.copy_(rgb[:ne])
    depth_np, rgb_np = depth_h.numpy(), rgb_h.numpy()
    g = new_map(args, rank, world, local, max_num_triangles=4_000_000 if world > 1 else 1)
    stream = torch.cuda.ExternalStream(g.cudaStream(), device=dev)
    # N > 1: depth and colour travel in ONE buffer, one NCCL broadcast per frame; two buffers, so that
    # the ingest + broadcast of frame k+1 (torch's stream) overlaps the kernels of frame k (the handle's
    # stream). Ordering is by events only; the host never waits inside the loop.
    bcast = [torch.empty(7 * P, dtype=torch.uint8, device=dev) for _ in range(2)] if world > 1 else None
    bcast_d = [b[: 4 * P].view(torch.float32).view(args.height, args.width) for b in bcast] if world > 1 else None
    bcast_c = [b[4 * P :].view(args.height, args.width, 3) for b in bcast] if world > 1 else None
    scatter = world > 1 and args.height % world == 0 and os.environ.get("MRH_BENCH_INGEST", "scatter") != "broadcast"
    scatter_views = [sharding.scatter_views(bcast_d[b], bcast_c[b]) for b in range(2)] if scatter else None
    if scatter:  # this rank's band of every host frame, sliced once
        lo_r, hi_r = sharding.frame_row_band(rank, world, args.height)
        band_d = [depth_h[k][lo_r:hi_r] for k in range(ne)]
        band_c = [rgb_h[k][lo_r:hi_r] for k in range(ne)]
    torch_stream = torch.cuda.current_stream()
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    for e in ev_done:
        e.record(stream)
    if world > 1:
        g.setStatsPipeline(True)

    host_us = {"setters": 0.0, "compute": 0.0, "read_result": 0.0}
    if world == 1:
        # streaming caller: page-locked frames are read by DMA while the previous frame's kernel runs
        # (the frames of this pass are never modified), and every frame's counters reach the host two
        # frames late instead of draining the device after each compute()
        g.setIngestMode(2)
        g.setStatsPipeline(True)
    e2e_state = {"n": 0}

    def step_host(k):
        if world == 1:
            # where the host's wall time goes (reported as e2e.host_us_per_step)
            c0 = time.perf_counter()
            g.setCurrPose(*poses[k])
            g.setDepthImage(depth_np[k])
            g.setRGBImage(rgb_np[k])
            c1 = time.perf_counter()
            g.compute()
            c2 = time.perf_counter()
            # D2H read of a step's result: the counters of the frame submitted STATS_LAG steps ago (each
            # frame's copy is enqueued behind its kernel); the last STATS_LAG frames' are read after the
            # loop. With a lag of 2 the host never waits for the kernel it has just queued, so the
            # upload of the next frame is submitted while the previous kernel is still running.
            st = g.getStatsPipelined(STATS_LAG) if e2e_state["n"] >= STATS_LAG else None
            e2e_state["n"] += 1
            c3 = time.perf_counter()
            host_us["setters"] += c1 - c0
            host_us["compute"] += c2 - c1
            host_us["read_result"] += c3 - c2
            return st
        # N > 1: rank 0 ingests the frame from its host buffers; the others receive it over NVLink.
        # The copy + broadcast run on torch's stream; the handle's stream is ordered after them
        # (and the broadcast that reuses this buffer, two frames on, after this frame's kernels).
        b = e2e_state["n"] & 1
        g.setCurrPose(*poses[k])
        torch_stream.wait_event(ev_done[b])
        if scatter:
            # every rank uploads its band of rows over its own PCIe link; an in-place all-gather over
            # NVLink completes the frame everywhere (sharding.scatter_ingest_frame)
            sharding.scatter_ingest_frame(bcast_d[b], bcast_c[b], band_d[k], band_c[k], views=scatter_views[b])
        else:
            if rank == 0:
                bcast_d[b].copy_(depth_h[k], non_blocking=True)
                bcast_c[b].copy_(rgb_h[k], non_blocking=True)
            dist.broadcast(bcast[b], 0)
        ev_ready[b].record()
        stream.wait_event(ev_ready[b])
        g.setDepthImageDevice(bcast_d[b].data_ptr(), args.height, args.width)
        g.setRGBImageDevice(bcast_c[b].data_ptr(), args.height, args.width)
        run_frame(g)
        ev_done[b].record(stream)
        st = g.getStatsPipelined(STATS_LAG) if e2e_state["n"] >= STATS_LAG else None  # an earlier frame's counters: no drain
        e2e_state["n"] += 1
        return st

    for k in range(W):
        step_host(k)
    barrier()
    if scatter and W > 0:
        # the band upload + all-gather must have delivered the whole frame to every rank (checked once, untimed)
        kb = (e2e_state["n"] - 1) & 1
        assert torch.equal(bcast_d[kb].cpu(), depth_h[W - 1]) and torch.equal(bcast_c[kb].cpu(), rgb_h[W - 1]), "scatter ingest delivered a different frame"
    host_us = {k: 0.0 for k in host_us}
    windows.append([time.time(), None])
    prof_range(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for i in range(Ke):
        step_host(W + i)
    for lag in range(STATS_LAG - 1, -1, -1):  # the results of the last steps (waits for those frames only)
        last_stats = g.getStatsPipelined(lag)
    assert last_stats["frames"] >= Ke, last_stats
    e1.record(stream)
    g.synchronize()
    barrier()
    wall_e2e = time.perf_counter() - t0
    prof_range(False)
    windows[-1][1] = time.time()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_e2e * 1e3))
    g_e2e = g  # NCCL enqueued work on this handle's stream: destroy it only after the process group
    if world == 1:
        g.close()

    # ---------------- pass B': the same from PAGEABLE numpy arrays (what rgbd_runner.py hands over) --
    e2e_pageable = None
    if world == 1:
        Kq = min(K, 300)
        depth_pg = [np.array(depth_np[W + i]) for i in range(Kq)]  # fresh pageable copies
        rgb_pg = [np.array(rgb_np[W + i]) for i in range(Kq)]
        g = new_map(args, rank, world, local)
        for k in range(W):
            g.setCurrPose(*poses[k])
            g.setDepthImage(depth_np[k])
            g.setRGBImage(rgb_np[k])
            g.compute()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.setStatsPipeline(True)
        for i in range(Kq):
            g.setCurrPose(*poses[W + i])
            g.setDepthImage(depth_pg[i])
            g.setRGBImage(rgb_pg[i])
            g.compute()
            if i >= STATS_LAG:
                g.getStatsPipelined(STATS_LAG)
        for lag in range(min(STATS_LAG, Kq) - 1, -1, -1):
            g.getStatsPipelined(lag)
        dt = time.perf_counter() - t0
        e2e_pageable = {"value": Kq / dt, "unit": "frames/s", "ms_per_step": 1e3 * dt / Kq, "steps": Kq, "what": "inputs in pageable host memory (what rgbd_runner.py hands over): copied into pinned staging inside the setters (default ingest mode), counters read two frames late"}
        g.close()

    # ---------------- N > 1 only: (i) one independent stream per GPU, (ii) sharded meshing ------------
    extra_multi = {}
    if world > 1:
        # (i) replicas: every rank maps its own copy of the stream, unsharded, no collective - the
        # aggregate a multi-session server gets from the box (weak scaling of the same hot path)
        g = new_map(args, 0, 1, local)
        stream = torch.cuda.ExternalStream(g.cudaStream(), device=dev)
        for k in range(W):
            step_device(k)
        g.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        r0.record(stream)
        for i in range(K):
            step_device(W + i)
        r1.record(stream)
        g.synchronize()
        barrier()
        ms_rep = max_over_ranks(r0.elapsed_time(r1))
        g.close()
        extra_multi["replica_streams"] = {"value": world * K / (ms_rep * 1e-3), "unit": "frames/s", "scaling": "weak", "what": "one unsharded stream per GPU, L2 warm, aggregate over ranks",
                                          "per_gpu": K / (ms_rep * 1e-3)}
        # `bench.py --gpus 1` runs configs[1] (640x480); what ONE GPU does on THIS line's workload is the
        # per-GPU figure above (an unsharded map of the same stream), to be compared with stream_fps_l2_warm
        extra_multi["single_gpu_same_workload"] = {"value": K / (ms_rep * 1e-3), "unit": "frames/s", "what": "one GPU, unsharded map, the same 1280x960 stream, L2 as the stream leaves it (compare: stream_fps_l2_warm of this line)"}
        # (ii) meshing the sharded map of the e2e pass where it lies: boundary exchange (two NCCL
        # all-to-alls), marching cubes per rank, soup gather + weld on rank 0
        barrier()
        t0 = time.perf_counter()
        sharding.extract_mesh_sharded(g_e2e, None, dst=0)  # first call: NCCL sets up its all-to-all peer connections
        barrier()
        t1 = time.perf_counter()
        _, info = sharding.extract_mesh_sharded(g_e2e, None, dst=0)
        barrier()
        info["ms"] = max_over_ranks((time.perf_counter() - t1) * 1e3)
        info["ms_first_call"] = max_over_ranks((t1 - t0) * 1e3)
        extra_multi["sharded_mesh"] = info

    # ---------------- roofline pass: per-kernel CUDA-event times, L2 flushed ------------------------
    Kp = min(K, 200)
    g = new_map(args, rank, world, local)
    stream = torch.cuda.ExternalStream(g.cudaStream(), device=dev)
    for k in range(W):
        step_device(k)
    g.synchronize()
    g.resetStats()
    g.setProfiling(True)
    with torch.cuda.stream(stream):
        for i in range(Kp):
            flush.zero_()
            step_device(W + i)
    g.synchronize()
    kms, kn = g.kernelTimes()
    stp = g.getStats()
    g.setProfiling(False)
    g.close()
    sampler.stop()

    peak, peak_src = peaks()
    # profiling slot 0 = k_frame (the whole frame is one launch; on the every-n-th starve frame the GC
    # tail kernels follow it and are not part of this window)
    fused = os.environ.get("MRH_FRAME", "fused") != "split"  # the library's own switch (mrh_capi.cu)
    # algorithmic bytes per frame: 7 B per pixel (depth + colour read once) + 24 B per
    # visible / new block (its table entry) + 24 B per updated voxel (12 read + 12 written)
    bytes_frame = 7.0 * P + (24.0 * (stp["blocks_visible"] + stp["blocks_new"]) + 24.0 * stp["voxels_updated"]) / Kp
    if fused:
        names = ["k_frame"]
        per_kernel = {"k_frame": {"ms_per_launch": kms[0] / max(kn[0], 1), "launches": kn[0]}}
        algo = {"k_frame": bytes_frame}
    else:  # MRH_FRAME=split: the two-launch frame of round 1
        names = ["k_front", "k_integrate"]
        per_kernel = {"k_front": {"ms_per_launch": kms[0] / max(kn[0], 1), "launches": kn[0]}, "k_integrate": {"ms_per_launch": kms[2] / max(kn[2], 1), "launches": kn[2]}}
        algo = {"k_integrate": (24.0 * stp["voxels_updated"] + 32.0 * stp["blocks_visible"]) / Kp + 7.0 * P,
                "k_front": 4.0 * P + 128.0 * stp["blocks_new"] / Kp + 32.0 * stp["live_blocks"] + 32.0 * stp["blocks_visible"] / Kp}
    dom = max(names, key=lambda k: per_kernel[k]["ms_per_launch"])
    achieved = algo[dom] / (per_kernel[dom]["ms_per_launch"] * 1e-3) / 1e9 if per_kernel[dom]["ms_per_launch"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(dom)
        except Exception:
            traffic = None

    # ---------------- N = 1: BASELINE configs[2] (LiDAR) and configs[4] (50 M-voxel mesh) ---------------
    extras = {}
    if world == 1 and not args.no_extras:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        del depth, rgb, flush
        torch.cuda.empty_cache()
        try:
            from oracle_lib import ref_available

            with_ref = ref_available()
        except Exception:
            with_ref = False
        log = lambda msg: print(msg, file=sys.stderr, flush=True)
        try:
            import bench_lidar

            extras["lidar"] = bench_lidar.run(40, 5, with_ref)
            for key in ("sdf_var_threshold=0.0", "sdf_var_threshold=0.005"):
                row = extras["lidar"][key]
                row["roofline"] = {"bound": "hbm", "achieved": row["algorithmic_bytes_per_frame"] / (row["device_ms_per_frame"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s"}
                row["roofline"]["frac"] = row["roofline"]["achieved"] / peak
        except Exception as e:  # an extra must never take the headline line down with it
            extras["lidar"] = {"error": repr(e)}
        try:
            import bench_mesh

            extras["mesh_50M"] = bench_mesh.run(97657, 0.003, with_ref, log)
            mrow = extras["mesh_50M"]
            mrow["roofline"] = {"bound": "hbm", "kernel": "k_mc_blocks", "achieved": mrow["mc_kernel_gbs"], "peak": peak, "unit": "GB/s", "frac": mrow["mc_kernel_gbs"] / peak}
        except Exception as e:
            extras["mesh_50M"] = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(args, depth_np, rgb_np, poses, args.cpu_seconds)
        fps = K / (ms_flushed * 1e-3)
        line = {
            "metric": "rgbd_frames_per_sec",
            "value": fps,
            "unit": "frames/s",
            "n_gpus": world,
            "steps": K,
            "warmup": W,
            "ms_per_step": ms_flushed / K,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": workload(args),
                "num_sdf_blocks": NUM_SDF_BLOCKS,
                "hash_num_buckets": HASH_NUM_BUCKETS,
                "l2": "flushed before every step (256 MiB memset, excluded from the per-step CUDA-event window)",
                "host_placement": numa,
                "parallelism": "1 GPU" if world == 1 else f"map sharded by hash-bucket range over {world} GPUs; e2e: frame " + ("uploaded in row bands by all ranks + NCCL all-gather" if scatter else "uploaded by rank 0 + NCCL broadcast"),
            },
            "mvoxels_updated_per_sec": v_upd / (ms_flushed * 1e-3) / 1e6,
            "voxels_updated_per_frame": v_upd / K,
            "visible_blocks_per_frame": b_vis / K,
            "new_blocks_per_frame": b_new / K,
            "live_blocks_end": live,
            "stream_fps_l2_warm": K / (ms_stream * 1e-3),
            "value_kind": "device time: CUDA events around each compute(), device-resident inputs, L2 flushed before every step",
            "e2e": {"value": Ke / (ms_e2e * 1e-3), "unit": "frames/s", "steps": Ke, "h2d_bytes_per_step": P * 7, "d2h_bytes_per_step": COUNTERS_BYTES, "ms_per_step": ms_e2e / Ke, "host_us_per_step": {k: 1e6 * v / Ke for k, v in host_us.items()},
                    "what": "page-locked host frames in (mrh_set_ingest_mode 2: DMA overlaps the previous frame's kernel), counters of every frame read back two frames late (mrh_get_stats_pipelined)" if world == 1 else ("every rank uploads its band of rows from page-locked host memory over its own PCIe link, in-place NCCL all-gather over NVLink completes the frame on every rank (sharding.scatter_ingest_frame), alternating buffers" if scatter else "rank 0 ingests from page-locked host frames, one NCCL broadcast per frame into alternating buffers") + " (overlaps the previous frame's kernels), counters of every frame read back two frames late"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo[dom], "ms_per_launch": per_kernel[dom]["ms_per_launch"]},
            "kernels": per_kernel,
            "frame_algorithmic_bytes": 7.0 * P * world + (24.0 * (b_vis + b_new) + 24.0 * v_upd) / K,
            "clocks": sampler.summary([tuple(w) for w in windows]),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if e2e_pageable is not None:
            line["e2e_pageable"] = e2e_pageable
        line.update(extra_multi)
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        g_e2e.close()


if __name__ == "__main__":
    main()
