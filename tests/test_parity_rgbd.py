"""GPU parity of the RGB-D compute() path: mrhash_b200 (through the C ABI) vs the CPU oracle vs the
unmodified reference kernels (oracle/_ref) on the same seeded synthetic frames.

Contract (BASELINE.json north_star): bit-exact block sets / weights / colours, sdf within 1e-5 relative.
"""
import numpy as np
import pytest

from compare import compare_dumps
from oracle_lib import Oracle, RefCuda, ref_available

from mrhash_b200 import GeoWrapper, synth

pytestmark = pytest.mark.gpu

NUM_BLOCKS = 60000
NUM_BUCKETS = 30000


def make_all(params, width=640, height=480, with_ref=True, num_blocks=NUM_BLOCKS, num_buckets=NUM_BUCKETS):
    fx, fy, cx, cy = synth.intrinsics(width, height)
    ours = GeoWrapper(**params, num_sdf_blocks=num_blocks, hash_num_buckets=num_buckets, max_num_triangles=1)
    ours.setCamera(fx, fy, cx, cy, height, width, params["min_depth"], params["max_depth"], 0)
    orc = Oracle(params, num_blocks, num_buckets)
    orc.set_camera(fx, fy, cx, cy, height, width, params["min_depth"], params["max_depth"], 0)
    ref = None
    if with_ref and ref_available():
        ref = RefCuda(params, num_blocks, num_buckets)
        ref.set_camera(fx, fy, cx, cy, height, width, params["min_depth"], params["max_depth"], 0)
    return ours, orc, ref


def feed(ours, others, t, q, depth, rgb):
    ours.setCurrPose(t, q)
    ours.setDepthImage(depth)
    ours.setRGBImage(rgb)
    ours.compute()
    st = ours.getStats()  # a dropped block would be silent (the reference prints and goes on): every parity frame asserts none
    assert st["dropped_heap"] == 0 and st["dropped_table"] == 0, st
    T = ours.getCurrPose()
    assert np.array_equal(T, synth.quat_to_matrix_f32(t, q))
    for o in others:
        if o is not None:
            o.compute_rgbd(T, depth, rgb)


def report(tag, rep):
    print(f"[{tag}] " + ", ".join(f"{k}={v}" for k, v in rep.items()))


def test_single_frame_identity_pose():
    """BASELINE config 1: one 640x480 frame, identity pose, replica.cfg parameters."""
    params = dict(synth.REPLICA_PARAMS)
    ours, orc, ref = make_all(params)
    t, q, depth, rgb = synth.rgbd_frame(0, orbit=False)
    feed(ours, [orc, ref], t, q, depth, rgb)
    mine = ours.dumpState()
    st = ours.getStats()
    assert st["dropped_heap"] == 0 and st["dropped_table"] == 0
    assert len(mine[0]) > 500
    rep_o = compare_dumps(mine, orc.dump())
    report("ours-vs-oracle", rep_o)
    if ref is not None:
        rep_r = compare_dumps(mine, ref.dump())
        report("ours-vs-refcuda", rep_r)
        assert rep_r["ok"], rep_r
        assert st["heap_free"] == ref.heap_high_free()
    assert rep_o["ok"], rep_o
    assert st["heap_free"] == orc.heap_high_free()
    assert st["voxels_updated"] == orc.stats()["voxels_updated"]


@pytest.mark.parametrize("n_frames,n_gc", [(12, 100), (12, 5)])
def test_orbit_sequence(n_frames, n_gc):
    """Short S2 stream (orbiting camera, 3.6 deg/frame), with and without starve frames."""
    params = dict(synth.REPLICA_PARAMS)
    params["n_frames_invalidate_voxels"] = n_gc
    ours, orc, ref = make_all(params)
    for k in range(n_frames):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=100)
        feed(ours, [orc, ref], t, q, depth, rgb)
    mine = ours.dumpState()
    st = ours.getStats()
    assert st["dropped_heap"] == 0 and st["dropped_table"] == 0
    rep_o = compare_dumps(mine, orc.dump())
    report(f"orbit{n_frames}/gc{n_gc} ours-vs-oracle", rep_o)
    if ref is not None:
        rep_r = compare_dumps(mine, ref.dump())
        report(f"orbit{n_frames}/gc{n_gc} ours-vs-refcuda", rep_r)
        assert rep_r["ok"], rep_r
    assert rep_o["ok"], rep_o


def test_invariants_after_sequence():
    """The reference's own invariants (tests/test_hash_utils.cu:378-526 HeapSanityCheck,
    :192-304 AllocationDeletion): no duplicate keys, occupied + free == num_sdf_blocks."""
    params = dict(synth.REPLICA_PARAMS)
    ours, _, _ = make_all(params, with_ref=False)
    for k in range(6):
        t, q, depth, rgb = synth.rgbd_frame(k, n_frames=60)
        ours.setCurrPose(t, q)
        ours.setDepthImage(depth)
        ours.setRGBImage(rgb)
        ours.compute()
    entries, voxels = ours.dumpState()
    st = ours.getStats()
    keys = {tuple(e[:3]) for e in entries}
    assert len(keys) == len(entries)
    assert len(entries) + st["heap_free"] == NUM_BLOCKS
    assert len(set(entries[:, 4].tolist())) == len(entries)  # every pool block used once
    assert (entries[:, 4] % 512 == 0).all()


def test_spherical_camera_range_images():
    """The depth-image path with the spherical camera model (camera.cuh:97-103,146-160,183-199: a LiDAR
    range image fed through setDepthImage): cloud depth = ||inverseProjection||, projection through
    atan2f / asinf. Ours vs the reference kernels; the CPU oracle's libm differs from libdevice in the
    last ulp of the trigonometry, so it gets the tolerance of the parity contract instead of bit equality."""
    params = dict(synth.REPLICA_PARAMS)
    params.update(virtual_voxel_size=0.05, sdf_truncation=0.2, min_depth=0.3, max_depth=20.0, n_frames_invalidate_voxels=2)
    rows, cols = 64, 512
    K = (-cols / (2 * np.pi), -rows / (np.pi / 2), cols / 2, rows / 2)
    nb, nk = 60000, 30000
    ours = GeoWrapper(**params, num_sdf_blocks=nb, hash_num_buckets=nk, max_num_triangles=1)
    ours.setCamera(*K, rows, cols, params["min_depth"], params["max_depth"], 1)
    ref = None
    if ref_available():
        ref = RefCuda(params, nb, nk)
        ref.set_camera(*K, rows, cols, params["min_depth"], params["max_depth"], 1)
    rng = np.random.default_rng(4)
    az = np.linspace(0, 4 * np.pi, cols, dtype=np.float32)[None, :]
    el = np.linspace(-1, 1, rows, dtype=np.float32)[:, None]
    for k in range(4):
        depth = (3.0 + 0.8 * np.sin(az + 0.3 * k) * np.cos(2 * el) + 0.3 * el).astype(np.float32)
        depth[:, 100:120] = 0.0  # invalid sector
        depth[5:8, :] = 50.0  # beyond max_depth
        rgb = rng.integers(0, 256, size=(rows, cols, 3), dtype=np.uint8)
        t = np.array([0.05 * k, -0.02 * k, 0.01 * k])
        q = np.array([0.0, 0.0, np.sin(0.02 * k), np.cos(0.02 * k)])
        feed(ours, [ref], t, q, depth, rgb)
    mine = ours.dumpState()
    st = ours.getStats()
    assert len(mine[0]) > 1000 and st["voxels_updated"] > 100000 and st["dropped_heap"] == 0 and st["dropped_table"] == 0
    if ref is not None:
        rep = compare_dumps(mine, ref.dump())
        report("spherical ours-vs-reference", rep)
        assert rep["ok"] and rep["sdf_bitexact"] and rep["sum_squared_bitexact"], rep
        assert st["heap_free"] == ref.heap_high_free()
