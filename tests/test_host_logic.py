"""CPU: host-side pieces that need no device - synthetic streams, drop-in import path, bench plumbing."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

from mrhash_b200 import synth


def test_dropin_import_path():
    # apps/rgbd_runner.py:9 does `from mrhash.src.pygeowrapper import GeoWrapper`
    from mrhash.src.pygeowrapper import GeoWrapper as A
    from mrhash_b200 import GeoWrapper as B

    assert A is B
    import inspect

    sig = inspect.signature(A.__init__)
    # pygeowrapper.cpp:14-29, in order
    want = ["sdf_truncation", "sdf_truncation_scale", "integration_weight_sample", "virtual_voxel_size", "n_frames_invalidate_voxels", "voxel_extents_scale", "viewer_active", "marching_cubes_threshold", "min_weight_threshold", "min_depth", "max_depth", "gs_optimization_param_path", "sdf_var_threshold", "vertices_merging_threshold", "projective_sdf"]
    assert list(sig.parameters)[1 : 1 + len(want)] == want
    for m in ["getHashNumBuckets", "getNumSdfBlocks", "getCurrPose", "getVertices", "getFaces", "getColors", "setRGBImage", "setDepthImage", "setPointCloud", "setCamera", "setCurrPose", "setCameraInLidar", "compute", "extractMesh", "streamAllOut", "clearBuffers", "serializeData", "serializeGrid", "deserializeGrid", "GSSavePointCloud", "GSFinalOpt"]:
        assert callable(getattr(A, m)), m


def test_synthetic_rgbd_stream_is_seeded_and_plausible():
    t, q, d, rgb = synth.rgbd_frame(0, orbit=False)
    t2, q2, d2, rgb2 = synth.rgbd_frame(0, orbit=False)
    assert np.array_equal(d, d2) and np.array_equal(rgb, rgb2)
    assert d.dtype == np.float32 and rgb.dtype == np.uint8 and d.shape == (480, 640) and rgb.shape == (480, 640, 3)
    assert 0.7 < d.min() < 0.9 and d.max() == 2.0  # sphere in front, wall at z = 2
    assert np.allclose(d * synth.DEPTH_SCALE, np.round(d * synth.DEPTH_SCALE), atol=1e-2)  # Replica PNG quantisation
    # orbit: unit-radius circle, optical axis radially outward
    for k in (0, 250, 500):
        t, q, R = synth.orbit_pose(k, 1000)
        assert abs(np.linalg.norm(t) - 1) < 1e-6 and np.allclose(R[:, 2] * 1.0, t, atol=1e-6)
        assert np.allclose(synth.quat_to_matrix_f32(t, q)[:3, :3], R, atol=1e-6)
    # the torch renderer used by bench.py produces the same frame
    d3, c3 = synth.render_rgbd_torch(np.eye(3), np.zeros(3), device="cpu")
    assert np.array_equal(d3.numpy(), d2) and np.array_equal(c3.numpy(), rgb2)


def test_synthetic_lidar_stream():
    T, pts = synth.lidar_frame(0)
    assert pts.dtype == np.float32 and pts.shape[1] == 3 and 100000 < len(pts) <= 128 * 1024
    r = np.linalg.norm(pts, axis=1)
    assert r.min() > 0.2 and r.max() < 100.0
    assert np.array_equal(T[:3, :3], np.eye(3, dtype=np.float32))


def test_bench_reference_arm_reports_unavailable_without_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    import torch

    if not torch.cuda.is_available():
        assert "unavailable" in line
