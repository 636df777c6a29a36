"""The small seeded cases shared by tests/golden/make_golden.py (reference kernels, GPU) and
tests/test_oracle_golden.py (CPU oracle): same inputs, same digests."""
import zlib

import numpy as np

from mrhash_b200 import synth


def _p(base, **kw):
    p = dict(base)
    p.update(kw)
    return p


LIDAR_ROWS, LIDAR_COLS = 32, 256
LIDAR_K = (-LIDAR_COLS / (2 * np.pi), -LIDAR_ROWS / (np.pi / 2), LIDAR_COLS / 2, LIDAR_ROWS / 2)

CASES = {
    # BASELINE config 1 in miniature: one frame, identity pose, replica.cfg parameters
    "rgbd_single": dict(kind="rgbd", params=_p(synth.REPLICA_PARAMS), width=160, height=120, frames=[(0, False)], num_blocks=20000, num_buckets=10000),
    # orbit with garbage collection and a starve frame (n = 3 -> starve on frames 3 and 6)
    "rgbd_orbit_gc": dict(kind="rgbd", params=_p(synth.REPLICA_PARAMS, n_frames_invalidate_voxels=3), width=160, height=120, frames=[(k, True) for k in range(8)], orbit_frames=120, num_blocks=20000, num_buckets=10000, starve_ties=True),
    # variance-adaptive path (streamer_example.cfg threshold)
    "rgbd_variance": dict(kind="rgbd", params=_p(synth.REPLICA_PARAMS, sdf_var_threshold=1.0), width=160, height=120, frames=[(k, True) for k in range(5)], orbit_frames=400, noise=0.002, num_blocks=20000, num_buckets=10000),
    # marching cubes after 6 small-step frames
    "mesh": dict(kind="rgbd", params=_p(synth.REPLICA_PARAMS), width=160, height=120, frames=[(k, True) for k in range(6)], orbit_frames=2000, num_blocks=20000, num_buckets=10000, max_triangles=400000, mesh=True),
    # LiDAR, one frame: block set + voxels hit by exactly one point (race-free in the reference)
    "lidar_single": dict(kind="lidar", params=_p(synth.VBR_PARAMS), frames=[0], num_blocks=60000, num_buckets=30000, racy=True),
}


def run_case(case, impl):
    """impl: Oracle or RefCuda (same mini-interface). Returns (entries, voxels, triangles or None)."""
    p = case["params"]
    if case["kind"] == "rgbd":
        w, h = case["width"], case["height"]
        fx, fy, cx, cy = synth.intrinsics(w, h)
        impl.set_camera(fx, fy, cx, cy, h, w, p["min_depth"], p["max_depth"], 0)
        for k, orbit in case["frames"]:
            t, q, depth, rgb = synth.rgbd_frame(k, n_frames=case.get("orbit_frames", 1000), width=w, height=h, orbit=orbit, noise_sigma=case.get("noise", 0.0))
            impl.compute_rgbd(synth.quat_to_matrix_f32(t, q), depth, rgb)
    else:
        impl.set_camera(*LIDAR_K, LIDAR_ROWS, LIDAR_COLS, p["min_depth"], p["max_depth"], 1)
        for k in case["frames"]:
            T, pts = synth.lidar_frame(k, rows=LIDAR_ROWS, cols=LIDAR_COLS, noise_sigma=0.01)
            impl.compute_points(T, pts)
    entries, voxels = impl.dump()
    tris = None
    if case.get("mesh"):
        tris, n = impl.extract_triangles(case["max_triangles"])
        assert n == len(tris)
    return entries, voxels, tris


def block_digest(entries, voxels, case):
    """Per block: [sum of weights, crc32 of the sdf bits, crc32 of sum_squared bits, crc32 of rgb]
    over the block's meaningful voxels (64 for resolution 1)."""
    out = np.zeros((len(entries), 4), np.uint32)
    for i in range(len(entries)):
        v = voxels[i, : (64 if entries[i, 3] == 1 else 512)]
        out[i, 0] = int(v["weight"].astype(np.uint32).sum())
        out[i, 1] = zlib.crc32(v["sdf"].tobytes())
        out[i, 2] = zlib.crc32(v["sum_squared"].tobytes())
        out[i, 3] = zlib.crc32(np.stack([v["r"], v["g"], v["b"]], -1).tobytes())
    return out


def triangle_digest(tris):
    """Order-independent digest of a triangle soup [T,3,6]."""
    flat = np.ascontiguousarray(tris.reshape(len(tris), -1))
    order = np.lexsort(flat.T[::-1])
    s = flat[order]
    return dict(n_triangles=np.int64(len(tris)), tri_crc=np.uint32(zlib.crc32(s.tobytes())), tri_pos_sum=tris[:, :, :3].astype(np.float64).sum(axis=(0, 1)), tri_head=s[:64].copy())
