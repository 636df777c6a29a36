"""LiDAR point-cloud stream (BASELINE configs[2], scene S3 of SURVEY.md §8d): 128 beams x 1024 azimuths,
vbr.cfg parameters, sensor translating 1 m/frame; single resolution (sdf_var_threshold = 0) and the
variance-adaptive two-resolution path (0.005). Times compute() end to end (host points in, counters
read back) for mrhash_b200 and, with MRH_BENCH_REF=1, for the reference kernels (oracle/_ref) on the
same frames. Usage: python tools/bench_lidar.py [n_frames] [warmup]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from mrhash_b200 import GeoWrapper, synth


def run(n_frames=40, warm=5, with_ref=False):
    """LiDAR stream through the public API; returns the dict the bench line carries under "lidar"."""
    ROWS, COLS = synth.LIDAR_ROWS, synth.LIDAR_COLS
    K = (-COLS / (2 * np.pi), -ROWS / (np.pi / 2), COLS / 2, ROWS / 2)
    NUM_BLOCKS, NUM_BUCKETS = 400000, 200000
    frames = [synth.lidar_frame(k, noise_sigma=0.01) for k in range(warm + 2 * n_frames)]
    out = {"workload": "S3 LiDAR stream, 128 beams x 1024 azimuths, vbr.cfg parameters (BASELINE configs[2]); end to end: host points in, counters read back every frame",
           "points_per_frame": int(np.mean([len(p) for _, p in frames])), "frames": n_frames, "warmup": warm}
    for thr in (0.0, 0.005):
        p = dict(synth.VBR_PARAMS)
        p["sdf_var_threshold"] = thr
        g = GeoWrapper(**p, num_sdf_blocks=NUM_BLOCKS, hash_num_buckets=NUM_BUCKETS, max_num_triangles=1)
        g.setCamera(*K, ROWS, COLS, p["min_depth"], p["max_depth"], 1)
        for T, pts in frames[:warm]:
            g.setCurrPoseMatrix(T), g.setPointCloud(pts, False), g.compute()
        g.synchronize()
        g.resetStats()
        l0 = g.launchCount()
        # device time of the frames (CUDA events around compute(), summed)
        dev_ms = 0.0
        t0 = time.perf_counter()
        host = [0.0, 0.0, 0.0]
        for T, pts in frames[warm : warm + n_frames]:
            c0 = time.perf_counter()
            g.setCurrPoseMatrix(T), g.setPointCloud(pts, False)
            c1 = time.perf_counter()
            g.compute()
            c2 = time.perf_counter()
            st = g.getStats()
            c3 = time.perf_counter()
            dev_ms += g.lastComputeMs()
            host[0] += c1 - c0
            host[1] += c2 - c1
            host[2] += c3 - c2
        dt = time.perf_counter() - t0
        launches = g.launchCount() - l0
        # the next n_frames of the sweep, as a streaming caller would submit them: every frame's counters are
        # still read, two frames late (mrh_get_stats_pipelined), so the host never drains the device
        g.setStatsPipeline(True)
        g.synchronize()
        t0 = time.perf_counter()
        for i, (T, pts) in enumerate(frames[warm + n_frames :]):
            g.setCurrPoseMatrix(T), g.setPointCloud(pts, False)
            g.compute()
            if i >= 2:
                g.getStatsPipelined(2)
        for lag in (1, 0):
            g.getStatsPipelined(lag)
        dt_pipe = time.perf_counter() - t0
        g.setStatsPipeline(False)
        n_pts = out["points_per_frame"]
        upd = st["voxels_updated"] / n_frames
        # SURVEY 8(d)-style algorithmic bytes of a point frame: 12 B per point read, 24 B per voxel update
        algo = 12.0 * n_pts + 24.0 * upd
        row = {
            "frames_per_sec_e2e": n_frames / dt,
            "frames_per_sec_e2e_pipelined": n_frames / dt_pipe,
            "ms_per_frame": 1e3 * dt / n_frames,
            "device_ms_per_frame": dev_ms / n_frames,
            "mvoxel_updates_per_sec": st["voxels_updated"] / dt / 1e6,
            "voxel_updates_per_frame": upd,
            "algorithmic_bytes_per_frame": algo,
            "live_blocks_end": st["live_blocks"],
            "launches_per_frame": launches / n_frames,
            "host_us_per_frame": {"setters": 1e6 * host[0] / n_frames, "compute": 1e6 * host[1] / n_frames, "read_result": 1e6 * host[2] / n_frames},
            "dropped": [st["dropped_heap"], st["dropped_table"], st["dropped_updates"]],
        }
        g.close()
        if with_ref:
            from oracle_lib import RefCuda

            cwd = os.getcwd()
            os.makedirs("/tmp/mrh_ref_run", exist_ok=True)
            os.chdir("/tmp/mrh_ref_run")
            devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
            os.dup2(devnull, 1)
            try:
                r = RefCuda(p, NUM_BLOCKS, NUM_BUCKETS)
                r.set_camera(*K, ROWS, COLS, p["min_depth"], p["max_depth"], 1)
                for T, pts in frames[:warm]:
                    r.compute_points(T, pts)
                t0 = time.perf_counter()
                integ = 0.0
                for T, pts in frames[warm : warm + n_frames]:
                    r.compute_points(T, pts)
                    integ += r.last_integrate_ms()
                dtr = time.perf_counter() - t0
            finally:
                os.dup2(saved, 1)
                os.chdir(cwd)
            row["reference_frames_per_sec_e2e"] = n_frames / dtr
            row["reference_integrate_only_ms"] = integ / n_frames
        out[f"sdf_var_threshold={thr}"] = row
    return out


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 5, os.environ.get("MRH_BENCH_REF", "0") == "1")))
